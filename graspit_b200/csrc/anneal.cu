// Batched simulated annealing driver: K independent chains, one neighbour per chain
// per step, evaluated by the batched pose evaluation, accepted by the Metropolis rule -- all on the device.
//
// Follows SimAnn (reference src/EGPlanner/simAnn.cpp):
//   cooling              :218-227   T = T0 exp(-c k^(1/d))                     (Ingber schedule)
//   neighborDistribution :236-250   y = T ((1 + 1/T)^|2u-1| - 1) sign(u - 1/2)
//   variableNeighbor     :175-215   v = value + y * maxJump, redrawn until inside [min, max]; circular variables
//                                   wrap by their range; values within TINY (1e-7) of a limit snap to it; <= 100 draws
//   iterate              :96-158    neighbour at T(YC,YDIMS) * NBR_ADJ; accept when
//                                   P = exp((E_old - E_new) ERR_ADJ / T(HC,HDIMS)) > U, P = 1 if E_new < E_old;
//                                   the step counter advances only when a legal neighbour was found (FAIL otherwise)
// The reference draws from libc rand() seeded with time(); here every draw is Philox4x32-10 keyed by the seed with
// counter (chain, step, draw block), so a chain's trajectory does not depend on the batch layout or the GPU count.
#include "b2c_internal.h"

namespace b2c {

__host__ __device__ inline void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned out[4]) {
  for (int r = 0; r < 10; r++) {
    unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
    unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Draws {      // sequential uniforms in (0,1) of one (chain, step)
  unsigned chain, step, k0, k1, block, have;
  unsigned buf[4];
  __device__ Draws(unsigned c, unsigned s, unsigned long long seed) : chain(c), step(s), k0((unsigned)seed), k1((unsigned)(seed >> 32)), block(0), have(0) {}
  __device__ double next() {
    if (have == 0) { philox4x32_10(chain, step, block, 0u, k0, k1, buf); block++; have = 4; }
    unsigned x = buf[4 - have]; have--;
    return ((double)x + 0.5) * (1.0 / 4294967296.0);
  }
};

__device__ __forceinline__ double cooling(double t0, double c, int k, int d) { return t0 * exp(-c * pow((double)k, 1.0 / (double)d)); }

// one neighbour per chain
__global__ void anneal_propose_kernel(AnnealParams p) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.nchain) return;
  const int k = p.k[c];
  const double T = cooling(p.T0, p.YC, k, p.YDIMS) * p.NBR_ADJ;
  Draws rng((unsigned)(p.chain_offset + c), (unsigned)p.step, p.seed);
  const float *cur = p.cur + (size_t)c * p.nvar;
  float *out = p.prop + (size_t)c * p.nvar;
  const double base = 1.0 + 1.0 / T;
  for (int i = 0; i < p.nvar; i++) {
    const double vmin = p.vmin[i], vmax = p.vmax[i], jump = p.vjump[i], range = fabs(vmax - vmin);
    const double val = (double)cur[i];
    double v = vmax + 1.0;
    int loop = 0;
    while (v > vmax || v < vmin) {
      loop++;
      double u = rng.next();
      double a = fabs(2.0 * u - 1.0);
      double y = T * (pow(base, a) - 1.0);
      if (u < 0.5) y = -y;
      v = val + y * jump;
      if (p.circular[i]) {
        if (v > vmax) v -= range; else if (v < vmin) v += range;
      }
      if (v > vmax && v - vmax < 1.0e-7) v = vmax;
      if (v < vmin && v - vmin > -1.0e-7) v = vmin;
      if (loop == 100) break;
    }
    // the reference stores whatever the 100th draw gave; keep the state inside its box for the fp32 wire format
    float vf = (float)v;
    if (vf > (float)vmax) vf = (float)vmax;
    if (vf < (float)vmin) vf = (float)vmin;
    out[i] = vf;
  }
}

// Metropolis step on the evaluated neighbours
__global__ void anneal_accept_kernel(AnnealParams p) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.nchain) return;
  if (!p.legal[c]) return;                       // SimAnn::FAIL: nothing changes, the step counter stays
  const int k = p.k[c];
  const double T = cooling(p.T0, p.HC, k, p.HDIMS);
  const double e_old = p.ERR_ADJ * (double)p.cur_energy[c], e_new = p.ERR_ADJ * (double)p.energy[c];
  double P = e_new < e_old ? 1.0 : exp((e_old - e_new) / T);
  unsigned r[4];
  philox4x32_10((unsigned)(p.chain_offset + c), (unsigned)p.step, 0xffffffffu, 1u, (unsigned)p.seed, (unsigned)(p.seed >> 32), r);
  double U = ((double)r[0] + 0.5) * (1.0 / 4294967296.0);
  if (P > U) {
    const float *src = p.prop + (size_t)c * p.nvar;
    float *dst = p.cur + (size_t)c * p.nvar;
    for (int i = 0; i < p.nvar; i++) dst[i] = src[i];
    p.cur_energy[c] = p.energy[c];
    if (p.energy[c] < p.best_energy[c]) {
      float *b = p.best + (size_t)c * p.nvar;
      for (int i = 0; i < p.nvar; i++) b[i] = src[i];
      p.best_energy[c] = p.energy[c];
      // candidate for this context's best list (merged by anneal_toplist_kernel)
      if (p.cand && p.energy[c] < *p.top_thresh) {
        int at = atomicAdd(p.cand_count, 1);
        if (at < p.nchain) p.cand[at] = c;
      }
    }
    atomicAdd(&p.counters[0], 1);                // jumps
  }
  p.k[c] = k + 1;
  atomicAdd(&p.counters[1], 1);                  // legal neighbours
}

// This context's BEST_LIST_SIZE best chains (by best-so-far energy), kept incrementally: one warp, lane = list slot.
// Candidates are the chains whose best improved below the list's current worst entry since the last call.  A chain
// already listed is updated in place; otherwise it replaces the worst entry if it beats it.  Then the rows
// (state || energy) of the listed chains are written out for the all-gather.  Empty slots carry energy 3e38.
__device__ __forceinline__ void toplist_update(const AnnealParams &p, int k, float *rows) {
  const int lane = threadIdx.x & 31;
  int chain = lane < k ? p.top_chain[lane] : -2;
  float e = lane < k ? (chain >= 0 ? p.best_energy[chain] : 3.0e38f) : -3.0e38f;   // lanes >= k never count as worst
  int n = *p.cand_count;
  const bool overflow = n > p.nchain;      // the queue was not drained for many steps: rescan every chain
  if (overflow) n = p.nchain;
  for (int i = 0; i < n; i++) {
    const int c = overflow ? i : p.cand[i];
    const float ce = p.best_energy[c];
    const unsigned hit = __ballot_sync(0xffffffffu, chain == c);
    if (hit) {
      if (chain == c) e = ce;
      continue;
    }
    // worst slot (largest energy; empty slots are 3e38), lowest lane on ties
    float w = e;
    for (int o = 16; o > 0; o >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (ce < w) {
      const unsigned who = __ballot_sync(0xffffffffu, e == w);
      if (lane == __ffs(who) - 1) { chain = c; e = ce; }
    }
  }
  float w = e;
  for (int o = 16; o > 0; o >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o));
  if (lane < k) {
    p.top_chain[lane] = chain;
    float *r = rows + (size_t)lane * (p.nvar + 1);
    for (int i = 0; i < p.nvar; i++) r[i] = chain >= 0 ? p.best[(size_t)chain * p.nvar + i] : 0.f;
    r[p.nvar] = e;
  }
  if (lane == 0) { *p.top_thresh = w; *p.cand_count = 0; }
}

__global__ void anneal_toplist_kernel(AnnealParams p, int k, float *rows) { toplist_update(p, k, rows); }

// Fused best-list update + all-gather: warp 0 brings the context's best list up to date and leaves its rows in shared
// memory; then warp w stores them straight into rank w's ring slot (step % depth, this rank) over NVLink (peer-mapped
// memory, plain stores) under a sequence lock: tag < 0 while writing, tag = step + 1 when complete.  No collective, no
// rendezvous: a rank never waits for another one; readers take whatever step has landed (exchange_collect_kernel).
__global__ void __launch_bounds__(32 * EX_MAX_WORLD) anneal_publish_kernel(AnnealParams p, ExchangeDev x, float *rows_local) {
  extern __shared__ float srows[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nf = x.rows * x.row_floats;
  if (warp == 0) {
    toplist_update(p, x.rows, srows);
    __syncwarp();
    if (rows_local) for (int i = lane; i < nf; i += 32) rows_local[i] = srows[i];
  }
  __syncthreads();
  if (warp >= x.world) return;
  float *slot = x.ring[warp] + ((size_t)(p.step % x.depth) * x.world + x.rank) * x.slot_floats;
  volatile int *tag = reinterpret_cast<volatile int *>(slot + nf);
  if (lane == 0) { *tag = -(p.step + 1); __threadfence_system(); }
  __syncwarp();
  for (int i = lane; i < nf; i += 32) slot[i] = srows[i];
  __threadfence_system();
  __syncwarp();
  if (lane == 0) *tag = p.step + 1;
}

// Reader side: one warp per source rank scans that rank's `depth` slots in the LOCAL ring, takes the newest complete one
// (largest tag; re-checked after the copy: a slot that changed underneath is skipped) and copies its rows out.
// steps[src] = step + 1 of what was taken, 0 if nothing has arrived yet.
__global__ void __launch_bounds__(32 * EX_MAX_WORLD) exchange_collect_kernel(ExchangeDev x, float *out, int *steps) {
  const int lane = threadIdx.x & 31, src = threadIdx.x >> 5;
  if (src >= x.world) return;
  const int nf = x.rows * x.row_floats;
  const float *base = x.ring[x.rank];
  int taken = 0;
  for (int attempt = 0; attempt < x.depth && !taken; attempt++) {
    // newest complete slot not yet rejected
    int best = 0, bslot = -1;
    for (int d = 0; d < x.depth; d++) {
      const float *slot = base + ((size_t)d * x.world + src) * x.slot_floats;
      const int t = *reinterpret_cast<const volatile int *>(slot + nf);
      if (t > best) { best = t; bslot = d; }
    }
    if (bslot < 0) break;
    const float *slot = base + ((size_t)bslot * x.world + src) * x.slot_floats;
    for (int i = lane; i < nf; i += 32) out[(size_t)src * nf + i] = reinterpret_cast<const volatile float *>(slot)[i];
    __threadfence();
    const int t1 = *reinterpret_cast<const volatile int *>(slot + nf);
    const bool okay = __all_sync(0xffffffffu, t1 == best);
    if (okay) taken = best;
  }
  if (!taken) for (int i = lane; i < nf; i += 32) out[(size_t)src * nf + i] = (i % x.row_floats == x.row_floats - 1) ? 3.0e38f : 0.f;
  if (lane == 0) steps[src] = taken;
}

cudaError_t launch_anneal_toplist(const AnnealParams &p, int k, float *rows, cudaStream_t s) {
  anneal_toplist_kernel<<<1, 32, 0, s>>>(p, k, rows);
  return cudaGetLastError();
}

cudaError_t launch_anneal_publish(const AnnealParams &p, const ExchangeDev &x, float *rows_local, cudaStream_t s) {
  anneal_publish_kernel<<<1, 32 * (x.world < 1 ? 1 : x.world), sizeof(float) * (size_t)x.rows * x.row_floats, s>>>(p, x, rows_local);
  return cudaGetLastError();
}
cudaError_t launch_exchange_collect(const ExchangeDev &x, float *out, int *steps, cudaStream_t s) {
  exchange_collect_kernel<<<1, 32 * (x.world < 1 ? 1 : x.world), 0, s>>>(x, out, steps);
  return cudaGetLastError();
}

cudaError_t launch_anneal_propose(const AnnealParams &p, cudaStream_t s) {
  anneal_propose_kernel<<<(p.nchain + 127) / 128, 128, 0, s>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_anneal_accept(const AnnealParams &p, cudaStream_t s) {
  anneal_accept_kernel<<<(p.nchain + 127) / 128, 128, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace b2c
