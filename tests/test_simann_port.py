"""CPU: the annealing restatement (oracle/port/simann.py) against the reference's own SimAnn compiled verbatim
(oracle/ref_simann.cpp: same rand() stream, same trajectory), values worked out by hand from the reference's formulas
(src/EGPlanner/simAnn.cpp:218-250), the Philox known-answer vectors, and the host-side planner logic."""
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import pytest

from graspit_b200.planner import BestList, SearchVariables, SimAnnParams, shard, state_distance
from oracle.port import simann


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    r = simann.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = simann.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = simann.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_closed_forms():
    p = SimAnnParams.ANNEAL_DEFAULT()
    # T = T0 exp(-c k^(1/d)): k = 30000, d = 8 -> 30000^(1/8) = 3.62779..., T = 1e6 exp(-25.3945...)
    T = simann.cooling(p.T0, p.YC, 30000, p.YDIMS)
    assert abs(30000 ** 0.125 - 3.6278) < 1e-4
    assert abs(T - 1.0e6 * np.exp(-7.0 * 30000 ** 0.125)) < 1e-18 and 9.0e-6 < T < 9.7e-6
    # prob: better energy always jumps; worse energy exp((e_old - e_new) / T)
    assert simann.prob(2.0, 1.0, T) == 1.0
    assert abs(simann.prob(1.0e-6 * 10.0, 1.0e-6 * 12.0, 1.0e-6) - np.exp(-2.0)) < 1e-15
    # neighbour map: y = T((1 + 1/T)^|2u-1| - 1), sign from u < 1/2; u = 1 -> +1, u = 0 -> -1, u = 1/2 -> 0
    assert abs(simann.neighbor_distribution(0.01, np.array([1.0]))[0] - 1.0) < 1e-12
    assert abs(simann.neighbor_distribution(0.01, np.array([0.0]))[0] + 1.0) < 1e-12
    assert simann.neighbor_distribution(0.01, np.array([0.5]))[0] == 0.0
    y = simann.neighbor_distribution(0.5, np.array([0.75]))[0]          # 0.5 (3^0.5 - 1)
    assert abs(y - 0.5 * (np.sqrt(3.0) - 1.0)) < 1e-15


def test_proposals_stay_in_range_and_are_reproducible():
    sv = SearchVariables.axis_angle_eigen(2)
    assert sv.names[:6] == ["Tx", "Ty", "Tz", "theta", "phi", "alpha"] and list(sv.vjump[6:]) == [2.0, 2.0]
    assert sv.circular.tolist() == [0, 0, 0, 0, 1, 0, 0, 0]
    p = SimAnnParams.ANNEAL_DEFAULT()
    n = 512
    cur = np.tile(sv.default.astype(np.float32)[None, :], (n, 1))
    k = np.full(n, p.K0)
    a = simann.propose(cur, k, sv, p, step=3, seed=1234)
    b = simann.propose(cur, k, sv, p, step=3, seed=1234)
    assert (a == b).all()
    assert (a >= sv.vmin[None, :].astype(np.float32)).all() and (a <= sv.vmax[None, :].astype(np.float32)).all()
    # chains are independent of the batch they sit in (counter = global chain id)
    c = simann.propose(cur[100:200], k[100:200], sv, p, step=3, seed=1234, chain_offset=100)
    assert (c == a[100:200]).all()
    assert len({tuple(r) for r in a}) > n // 2
    # hotter chains move further
    hot = simann.propose(cur, np.zeros(n, dtype=np.int64), sv, p, step=3, seed=1234)
    assert np.abs(hot - cur).mean() > 3 * np.abs(a - cur).mean()


def test_accept_rule():
    p = SimAnnParams.ANNEAL_DEFAULT()
    n = 1000
    cur = np.zeros((n, 8), np.float32); prop = np.ones((n, 8), np.float32)
    ce = np.full(n, 50.0, np.float32); k = np.full(n, p.K0)
    legal = np.ones(n, np.uint8); legal[::4] = 0
    better = np.full(n, 40.0, np.float32)
    c2, e2, k2, j = simann.accept(cur, ce, k, prop, legal, better, p, step=0, seed=7)
    assert j[legal == 1].all() and not j[legal == 0].any()
    assert (k2 == p.K0 + legal.astype(np.int64)).all() and (e2[legal == 1] == 40.0).all() and (c2[legal == 0] == 0).all()
    # slightly worse energies jump with probability exp(dE * ERR_ADJ / T): at K0 that is exp(-0.1 * 1e-6 / 9.4e-6)
    worse = np.full(n, 50.1, np.float32)
    _, _, _, j = simann.accept(cur, ce, k, prop, np.ones(n, np.uint8), worse, p, step=1, seed=7)
    T = simann.cooling(p.T0, p.HC, p.K0, p.HDIMS)
    expect = np.exp(-(np.float32(50.1) - np.float32(50.0)) * 1e-6 / T)
    assert abs(j.mean() - expect) < 0.05


def test_best_list_rules():
    ref = (np.array([1.0, 0, 0, 0]), np.zeros(3))
    bl = BestList(ref, max_radius=100.0, size=3)
    s = lambda tx, al=1.0: np.array([tx, 0, 0, 0.5, 0.2, al, 0, 0], dtype=np.float32)
    assert bl.offer(s(0), 10.0)
    assert not bl.offer(s(5), 11.0)            # within 0.2 of a better state: rejected
    assert bl.offer(s(5), 9.0)                 # within 0.2 of a worse state: replaces it
    assert len(bl.states) == 1 and bl.energies == [9.0]
    assert bl.offer(s(50), 12.0) and bl.offer(s(100), 8.0)
    assert bl.energies == [8.0, 9.0, 12.0]
    assert not bl.offer(s(-80), 13.0)          # list full: must beat the worst
    assert bl.offer(s(-80), 10.0) and bl.energies == [8.0, 9.0, 10.0]
    # distance = max(|dt| / maxRadius, 0.5 |angle| / pi)
    a = (np.array([1.0, 0, 0, 0]), np.zeros(3)); b = (np.array([0.0, 1.0, 0, 0]), np.array([30.0, 0, 0]))
    assert abs(state_distance(a, b, 100.0) - 0.5) < 1e-12
    assert abs(state_distance(a, (a[0], np.array([30.0, 40.0, 0])), 100.0) - 0.5) < 1e-12


def test_shard_partition():
    for total, world in [(65536, 8), (10, 3), (7, 8)]:
        parts = [shard(total, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == total
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


# ---- the restatement against the reference's own SimAnn::iterate ---------------------------------------------------------
class _LibcDraws:
    """uniforms as the reference draws them: ((double) rand()) / RAND_MAX (simAnn.cpp:240,129)"""
    RAND_MAX = 2147483647

    def __init__(self, seed):
        self.libc = ctypes.CDLL("libc.so.6")
        self.libc.srand(ctypes.c_uint(seed))
        self.count = 0

    def one(self):
        self.count += 1
        return self.libc.rand() / self.RAND_MAX

    def next(self, mask):
        return np.array([self.one() if m else 0.5 for m in mask])


def _closed_form_energy(x, w, c, cut):
    e = 0.0
    for i in range(len(x)):
        d = float(x[i]) - c[i]
        e += w[i] * d * d
    legal = not (np.sin(3.0 * float(x[0])) * np.cos(2.0 * float(x[-1])) > cut)
    return legal, e


@pytest.mark.parametrize("preset,name,cut", [(0, "ANNEAL_DEFAULT", 0.3), (3, "ANNEAL_STRICT", 0.3), (4, "ANNEAL_ONLINE", -0.2),
                                             (2, "ANNEAL_MODIFIED", 0.6)])
def test_port_reproduces_the_reference_simann_trajectory(preset, name, cut):
    from oracle import ref
    if not os.path.exists(ref.SIMANN_LIB_PATH):
        pytest.skip("oracle/_ref/libgraspit_ref_simann.so not built")
    neg, iters, seed = 2, 400, 4242 + preset
    nvar = 6 + neg
    rng = np.random.default_rng(5 + preset)
    start = np.array([10.0, -20.0, 30.0, 1.0, 0.5, 2.0, 0.3, -0.7])
    w = rng.uniform(0.5, 2.0, nvar); c = np.array([40.0, 10.0, -15.0, 2.0, -1.0, 1.0, 1.0, -2.0])
    _, e0 = _closed_form_energy(start, w, c, cut)
    R = ref.simann_run(neg, seed, preset, iters, start, w, c, cut, e0)
    # the reference's own variable table is the one the product ships
    sv = SearchVariables.axis_angle_eigen(neg)
    assert np.allclose(R["table"][:, 0], sv.vmin) and np.allclose(R["table"][:, 1], sv.vmax)
    assert np.allclose(R["table"][:, 2], sv.vjump) and (R["table"][:, 3] == sv.circular).all()
    p = getattr(SimAnnParams, name)()
    variables = SimpleNamespace(vmin=R["table"][:, 0], vmax=R["table"][:, 1], vjump=R["table"][:, 2], circular=R["table"][:, 3].astype(int))
    draws = _LibcDraws(seed)
    cur = start[None, :].copy(); ce = np.array([e0]); k = np.array([p.K0])
    results = {0: 0, 1: 0, 2: 0}
    for it in range(iters):
        # SimAnn::iterate: up to 11 neighbours until one is legal (simAnn.cpp:106-116)
        legal, prop, e = False, None, 0.0
        for attempt in range(11):
            prop = simann.propose(cur, k, variables, p, step=0, seed=0, draws=draws, dtype=np.float64)
            legal, e = _closed_form_energy(prop[0], w, c, cut)
            if legal:
                break
        if not legal:
            r = 0
        else:
            U = np.array([draws.one()])
            cur, ce, k, j = simann.accept(cur, ce, k, prop, np.array([1], np.uint8), np.array([e]), p, step=0, seed=0, U=U)
            r = 1 if j[0] else 2
        results[r] += 1
        assert r == R["result"][it], (it, r, R["result"][it])
        assert int(k[0]) == R["step"][it]
        assert np.allclose(cur[0], R["state"][it], rtol=1e-9, atol=1e-9), (it, cur[0], R["state"][it])
        assert ce[0] == pytest.approx(R["energy"][it], rel=1e-9, abs=1e-9)
    # the run exercised every outcome the rule has
    assert results[1] > 3 and results[2] > 3, results
