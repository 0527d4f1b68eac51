// TEST: host-side check of the compact triangle-triangle distance against the straightforward form.
// nvcc -O2 -std=c++17 -I graspit_b200/csrc tests/host/geom_check.cu -o /tmp/geom_check && /tmp/geom_check
#include <cstdio>
#include <random>
#include "geom.cuh"
using namespace b2c;

int main() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> u(-1, 1);
  int bad = 0, n = 0, badf = 0, ndeg = 0;
  double worst = 0, worstf = 0;
  for (int it = 0; it < 400000; it++) {
    double sc = it % 3 == 0 ? 1.0 : (it % 3 == 1 ? 10.0 : 0.2);
    double off = it % 5 == 0 ? 0.3 : 2.5;
    d3 a[3], b[3];
    for (int k = 0; k < 3; k++) {
      a[k] = mk<double>(u(rng) * sc, u(rng) * sc, u(rng) * sc);
      b[k] = mk<double>(u(rng) * sc + off * sc, u(rng) * sc, u(rng) * sc);
    }
    if (it % 7 == 0) { b[1] = b[0] + (a[1] - a[0]) * 0.7; }            // parallel edges
    if (it % 11 == 0) { b[2].z = b[1].z = b[0].z = 1.5 * sc; a[0].z = a[1].z = a[2].z = 0; }   // parallel faces
    if (tri_tri_intersect_exact(a[0], a[1], a[2], b[0], b[1], b[2])) continue;
    n++;
    d3 pa, pb, qa, qb;
    double d_ref = tri_tri_dist2<double>(a[0], a[1], a[2], b[0], b[1], b[2], pa, pb);
    bool deg;
    double d_new = tri_tri_dist2_compact<double>(a[0], a[1], a[2], b[0], b[1], b[2], qa, qb, deg);
    if (deg) { ndeg++; d_new = tri_tri_dist2<double>(a[0], a[1], a[2], b[0], b[1], b[2], qa, qb); }
    double rel = fabs(sqrt(d_new) - sqrt(d_ref)) / fmax(sqrt(d_ref), 1e-12);
    if (rel > worst) worst = rel;
    if (rel > 1e-9) { if (bad < 5) printf("double mismatch: %.12g vs %.12g\n", sqrt(d_ref), sqrt(d_new)); bad++; }
    // the reported points realise the reported distance and lie on the triangles' planes
    d3 dv = qa - qb;
    if (fabs(sqrt(dot(dv, dv)) - sqrt(d_new)) > 1e-9 * (1 + sqrt(d_new))) { if (bad < 5) printf("points do not realise the distance\n"); bad++; }
    // fp32, re-centred on a[0] as the kernels do
    f3 fa[3], fb[3];
    for (int k = 0; k < 3; k++) { fa[k] = tof(a[k] - a[0]); fb[k] = tof(b[k] - a[0]); }
    f3 xa, xb;
    float df = tri_tri_dist2_compact<float>(fa[0], fa[1], fa[2], fb[0], fb[1], fb[2], xa, xb, deg);
    if (deg) df = tri_tri_dist2<float>(fa[0], fa[1], fa[2], fb[0], fb[1], fb[2], xa, xb);
    double relf = fabs(sqrt((double)df) - sqrt(d_ref)) / fmax(sqrt(d_ref), 1e-12);
    double scale = 0;
    for (int k = 0; k < 3; k++) { scale = fmax(scale, fmax(fabs(fb[k].x), fmax(fabs(fb[k].y), fabs(fb[k].z)))); scale = fmax(scale, fmax(fabs(fa[k].x), fmax(fabs(fa[k].y), fabs(fa[k].z)))); }
    bool tiny = d_ref < 1e-4 * scale * scale;       // handled by the fp64 refinement pass in the kernels
    if (!tiny) { if (relf > worstf) worstf = relf; if (relf > 2e-5) { if (badf < 5) printf("float mismatch: %.9g vs %.9g (scale %g)\n", sqrt(d_ref), sqrt((double)df), scale); badf++; } }
  }
  printf("%d degenerate-edge pairs; ", ndeg); printf("%d pairs, worst rel (double) %.3g, worst rel (float, non-tiny) %.3g, bad %d / %d\n", n, worst, worstf, bad, badf);
  return (bad || badf) ? 1 : 0;
}
