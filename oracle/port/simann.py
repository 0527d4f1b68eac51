"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp64) of the reference's simulated annealing step for a batch
of independent chains.  Follows src/EGPlanner/simAnn.cpp: cooling :218-227, prob :229-234, neighborDistribution
:236-250, variableNeighbor :175-215 (no target state), iterate :96-158.

Parity status: PINNED.  The reference draws from libc rand() seeded with time(NULL) (simAnn.cpp:85); the product's
stream is Philox4x32-10 (counter = (chain, step, draw block), shared with the CUDA kernels), so product trajectories are
its own by design.  The restatement itself, though, takes its uniforms from any source (`draws=`, `U=`) and runs in
fp64 on request (`dtype=`): fed libc's rand() stream it reproduces, call for call, the trajectory of the reference's
own SimAnn::iterate compiled verbatim (oracle/ref_simann.cpp, tests/test_simann_port.py::test_port_reproduces_*).
"""
from __future__ import annotations

import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised Philox4x32-10; all inputs broadcastable uint32 arrays; returns 4 uint32 arrays"""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & M32 for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint64(int(k0) & 0xFFFFFFFF)
    k1 = np.uint64(int(k1) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M32, p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1, c2, c3


def _to_unit(x):
    return (x.astype(np.float64) + 0.5) / 4294967296.0


class ChainDraws:
    """sequential uniforms of (chain, step): draw j = word j%4 of block j//4"""

    def __init__(self, chains, step, seed):
        self.chains = np.asarray(chains, dtype=np.uint64)
        self.step, self.seed = step, int(seed)
        self.idx = np.zeros(len(self.chains), dtype=np.int64)

    def next(self, mask):
        """one uniform for every chain in mask (others get 0.5 and do not advance)"""
        out = np.full(len(self.chains), 0.5)
        if mask.any():
            j = self.idx[mask]
            w = philox4x32_10(self.chains[mask], np.uint64(self.step), (j // 4).astype(np.uint64), np.uint64(0),
                              self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF)
            words = np.stack(w, 1)
            out[mask] = _to_unit(words[np.arange(len(j)), j % 4])
            self.idx[mask] += 1
        return out


def cooling(t0, c, k, d):
    return t0 * np.exp(-c * np.power(np.asarray(k, dtype=np.float64), 1.0 / d))


def prob(e_old, e_new, T):
    with np.errstate(over="ignore"):
        return np.where(e_new < e_old, 1.0, np.exp(np.minimum((e_old - e_new) / T, 0.0)))


def neighbor_distribution(T, u):
    v = np.abs(2.0 * u - 1.0)
    y = T * (np.power(1.0 + 1.0 / T, v) - 1.0)
    return np.where(u < 0.5, -y, y)


def propose(cur, k, variables, params, step, seed, chain_offset=0, draws=None, dtype=np.float32):
    """one neighbour per chain (variableNeighbor).  cur: (n, nvar) states; k: (n,) per-chain step counters.
    draws: an object with next(mask) -> uniforms (default: the product's Philox stream); dtype: float32 as the engine
    stores states, float64 to follow the reference's doubles."""
    cur = np.asarray(cur, dtype=dtype)
    n, nvar = cur.shape
    T = cooling(params.T0, params.YC, k, params.YDIMS) * params.NBR_ADJ
    rng = draws if draws is not None else ChainDraws(np.arange(n) + chain_offset, step, seed)
    out = np.zeros((n, nvar), dtype=dtype)
    for i in range(nvar):
        vmin, vmax, jump = float(variables.vmin[i]), float(variables.vmax[i]), float(variables.vjump[i])
        rng_range = abs(vmax - vmin)
        val = cur[:, i].astype(np.float64)
        v = np.full(n, vmax + 1.0)
        for _ in range(100):
            todo = (v > vmax) | (v < vmin)
            if not todo.any():
                break
            u = rng.next(todo)
            cand = val + neighbor_distribution(T, u) * jump
            if variables.circular[i]:
                cand = np.where(cand > vmax, cand - rng_range, np.where(cand < vmin, cand + rng_range, cand))
            cand = np.where((cand > vmax) & (cand - vmax < 1.0e-7), vmax, cand)
            cand = np.where((cand < vmin) & (cand - vmin > -1.0e-7), vmin, cand)
            v = np.where(todo, cand, v)
        vf = v.astype(dtype)
        out[:, i] = np.minimum(np.maximum(vf, dtype(vmin)), dtype(vmax))
    return out


def accept(cur, cur_energy, k, prop, legal, energy, params, step, seed, chain_offset=0, U=None):
    """Metropolis rule of SimAnn::iterate; returns (cur, cur_energy, k, jumped).  U: the uniforms to compare P with
    (default: the product's Philox stream)"""
    n = len(k)
    T = cooling(params.T0, params.HC, k, params.HDIMS)
    P = prob(params.ERR_ADJ * cur_energy.astype(np.float64), params.ERR_ADJ * energy.astype(np.float64), T)
    if U is None:
        w = philox4x32_10(np.arange(n, dtype=np.uint64) + np.uint64(chain_offset), np.uint64(step), np.uint64(0xFFFFFFFF), np.uint64(1),
                          np.uint64(int(seed) & 0xFFFFFFFF), np.uint64((int(seed) >> 32) & 0xFFFFFFFF))
        U = _to_unit(w[0])
    ok = legal.astype(bool)
    jump = ok & (P > U)
    cur = np.where(jump[:, None], prop, cur)
    cur_energy = np.where(jump, energy, cur_energy)
    k = np.where(ok, k + 1, k)
    return cur, cur_energy, k, jump
