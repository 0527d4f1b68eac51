// The arithmetic of the reference's contact-candidate visit order, in ONE source shared by the host replay
// (contacts_host.cpp, obb_visit_order) and the device kernel that computes per-candidate visit keys (contact_order.cu):
// both sides execute the same IEEE double operations in the same order (the .cu is compiled with -fmad=false, g++ emits no
// fused multiply-adds for generic x86-64), so a decision `d1 < d0` falls the same way wherever it is taken.
//   bboxDistanceSq(bb1, bb2, tran2To1)      reference include/graspit/bbox_inl.h:269-413 (same operation order per axis)
//   split rule, first-child rule            reference src/Collision/Graspit/collisionAlgorithms.cpp:60-101
//   mTran2To1 = T1^-1 % T2                  reference src/Collision/Graspit/collisionAlgorithms.cpp:149-156
#pragma once
#ifdef __CUDACC__
#define B2C_HD __host__ __device__
#else
#define B2C_HD
#endif

namespace b2c {

// what the order rule reads of one node of the reference-style OBB tree (plain data: uploaded to the device as it is)
struct OrderNode {
  double Rq[3][3];            // box axes after the round trip through the unit quaternion transf stores (BoundingBox::setTran)
  double center[3];
  double half[3];
  int child0, child1;         // -1: leaf
  int leaf_lo, leaf_hi;       // leaves below this node, as a range of depth-first leaf positions
};

struct OrderXf {              // model 2 -> model 1
  double R[3][3];
  double t[3];
};

// model 2 -> model 1 from two world transforms in Xf layout (row-major rotation [0..8], translation [9..11])
B2C_HD inline void order_rel_xf(const double *X1, const double *X2, OrderXf &x) {
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) x.R[i][j] = X1[i] * X2[j] + X1[3 + i] * X2[3 + j] + X1[6 + i] * X2[6 + j];
    x.t[i] = X1[i] * (X2[9] - X1[9]) + X1[3 + i] * (X2[10] - X1[10]) + X1[6 + i] * (X2[11] - X1[11]);
  }
}

B2C_HD inline double order_box_volume(const OrderNode &n) { return n.half[0] * n.half[1] * n.half[2]; }

B2C_HD inline double order_abs(double v) { return v < 0.0 ? -v : v; }
B2C_HD inline double order_max(double a, double b) { return a < b ? b : a; }     // std::max(a, b)

B2C_HD inline void order_matmul(const double a[3][3], const double b[3][3], double r[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += a[i][k] * b[k][j];
      r[i][j] = s;
    }
}

// bboxDistanceSq(bb1, bb2, tran2To1); -1 for boxes the separating-axis bounds cannot tell apart
#ifndef B2C_ORDER_DIST_ATTR
#define B2C_ORDER_DIST_ATTR
#endif
B2C_HD B2C_ORDER_DIST_ATTR inline double order_distance_sq(const OrderNode &n1, const OrderNode &n2, const OrderXf &x) {
  // BtoA = bb1.TranInv % tran2To1 % bb2.Tran
  double R1t[3][3], M[3][3], B[3][3], Bf[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R1t[i][j] = n1.Rq[j][i];
  order_matmul(R1t, x.R, M);
  order_matmul(M, n2.Rq, B);
  double c2[3], dc[3], T[3];
  for (int i = 0; i < 3; i++) c2[i] = (x.R[i][0] * n2.center[0] + x.R[i][1] * n2.center[1] + x.R[i][2] * n2.center[2]) + x.t[i];
  for (int i = 0; i < 3; i++) dc[i] = c2[i] - n1.center[i];
  for (int i = 0; i < 3; i++) T[i] = R1t[i][0] * dc[0] + R1t[i][1] * dc[1] + R1t[i][2] * dc[2];
  const double *a = n1.half, *b = n2.half;
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) Bf[i][k] = order_abs(B[i][k]) + 1e-6;
  double ga[3], gb[3];
  for (int i = 0; i < 3; i++) {
    double d = order_abs(T[i]) - (a[i] + b[0] * Bf[i][0] + b[1] * Bf[i][1] + b[2] * Bf[i][2]);
    d = order_max(0.0, d);
    ga[i] = d * d;
  }
  for (int j = 0; j < 3; j++) {
    double s = T[0] * B[0][j] + T[1] * B[1][j] + T[2] * B[2][j];
    double d = order_abs(s) - (b[j] + a[0] * Bf[0][j] + a[1] * Bf[1][j] + a[2] * Bf[2][j]);
    d = order_max(0.0, d);
    gb[j] = d * d;
  }
  double maxD = order_max(ga[0] + ga[1] + ga[2], gb[0] + gb[1] + gb[2]);
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    const int alo = i1 < i2 ? i1 : i2, ahi = i1 < i2 ? i2 : i1;
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const int blo = j1 < j2 ? j1 : j2, bhi = j1 < j2 ? j2 : j1;
      double s = T[i2] * B[i1][j] - T[i1] * B[i2][j];
      // a[alo] pairs with the row of the OTHER index and vice versa, likewise for b (bbox_inl.h:341-409)
      double d = order_abs(s) - (a[alo] * Bf[ahi][j] + a[ahi] * Bf[alo][j] + b[blo] * Bf[i][bhi] + b[bhi] * Bf[i][blo]);
      d = order_max(0.0, d);
      d *= d;
      maxD = order_max(maxD, order_max(ga[i] + d, gb[j] + d));
    }
  }
  if (maxD == 0.0) return -1;
  return maxD;
}

// One step of the reference's dual recursion below the node pair (N1, N2), for a candidate whose triangles sit at leaf
// positions pa / pb: which node is split, which child holds the candidate, and whether that child is visited first.
//   returns false when both nodes are leaves (the walk is over)
//   range (optional) = {amin, amax, bmin, bmax}: leaf positions spanned by the candidates whose order is being decided;
//   where they all sit under one child of the split node nobody is on the other side, the bit cannot matter, and the two
//   box distances are not evaluated
// The step in two halves, so that a kernel can run the cheap half of many candidates up to their next deciding level and
// then evaluate the box distances of all of them together.
struct OrderPick {
  int split1;              // 1: model 1's node is split (nodeSwap == false)
  int c0, c1;              // its children
  int mine;                // the child holding the candidate
};
//   0: both nodes are leaves (the walk is over);  1: the level decides nothing for this group;  2: the level decides
B2C_HD inline int order_pick(const OrderNode *A, const OrderNode *B, int n1, int n2, int pa, int pb, const int *range, OrderPick &k) {
  const OrderNode &N1 = A[n1], &N2 = B[n2];
  const bool l1 = N1.child0 < 0, l2 = N2.child0 < 0;
  if (l1 && l2) return 0;
  bool split1;
  if (l1) split1 = false;
  else if (l2) split1 = true;
  else split1 = order_box_volume(N1) > order_box_volume(N2);
  const OrderNode &R = split1 ? N1 : N2;
  const OrderNode *T = split1 ? A : B;
  const OrderNode &C0 = T[R.child0];
  const int pos = split1 ? pa : pb;
  k.split1 = split1 ? 1 : 0;
  k.c0 = R.child0; k.c1 = R.child1;
  k.mine = (pos >= C0.leaf_lo && pos < C0.leaf_hi) ? 0 : 1;
  if (range) {
    const int pmin = split1 ? range[0] : range[2], pmax = split1 ? range[1] : range[3];
    if ((pmin >= C0.leaf_lo && pmax < C0.leaf_hi) || pmin >= C0.leaf_hi || pmax < C0.leaf_lo) return 1;
  }
  return 2;
}
//   which child the reference visits first
B2C_HD inline int order_first(const OrderNode *A, const OrderNode *B, const OrderXf &x, int n1, int n2, const OrderPick &k) {
  double d0, d1;
  if (k.split1) { d0 = order_distance_sq(A[k.c0], B[n2], x); d1 = order_distance_sq(A[k.c1], B[n2], x); }
  else { d0 = order_distance_sq(A[n1], B[k.c0], x); d1 = order_distance_sq(A[n1], B[k.c1], x); }
  return d1 < d0 ? 1 : 0;
}
B2C_HD inline void order_advance(int &n1, int &n2, const OrderPick &k) {
  if (k.split1) n1 = k.mine ? k.c1 : k.c0; else n2 = k.mine ? k.c1 : k.c0;
}

B2C_HD inline bool order_step(const OrderNode *A, const OrderNode *B, const OrderXf &x, int &n1, int &n2, int pa, int pb, int *second,
                              const int *range = nullptr) {
  OrderPick k;
  const int st = order_pick(A, B, n1, n2, pa, pb, range, k);
  if (st == 0) return false;
  *second = st == 2 && k.mine != order_first(A, B, x, n1, n2, k) ? 1 : 0;
  order_advance(n1, n2, k);
  return true;
}

} // namespace b2c
