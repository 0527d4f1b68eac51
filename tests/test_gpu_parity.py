"""Parity of the CUDA path (through the C ABI) with the oracle = the reference's own GraspitCollision.

Tolerances are BASELINE.json's: collision / contact pair sets bit-exact (outside 1e-5 mm of tangency),
distances and energies within 1e-4 relative (fp32)."""
import numpy as np
import pytest

from graspit_b200.engine import (ALL_COLLISIONS, B2C_INPUT_AA_EIGEN, FAST_COLLISION, B2CError, Engine, SceneBinding)
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from tests.common import OracleScene, aa_to_raw_states, object_body, sample_aa_states

pytestmark = pytest.mark.gpu
REL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    osc = OracleScene(sc, nthreads=min(8, ref.hardware_threads()))
    pos, dofs = sample_aa_states(sc, 1000, seed=1234)
    raw = aa_to_raw_states(sc, pos, dofs)
    return dict(sc=sc, eng=eng, sb=sb, osc=osc, raw=raw, pos=pos, dofs=dofs, obj=object_body(sc), rid=sb.robot_ids[0])


def relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-9)


def test_engine_loaded_native_library(ctx):
    assert ctx["eng"].L._name.endswith("libb2c.so")
    assert len(ctx["eng"].robot_pairs(ctx["rid"])) == 35          # 9 hand-vs-mug + 26 self pairs
    order = [ctx["sb"].body_ids[b] for b in ctx["sc"].robots[0].body_list()]
    assert ctx["eng"].robot_bodies(ctx["rid"]) == order


def test_config1_fk_legal_energy_preset(ctx):
    """config 1: 1000 poses, FK + allCollisions(FAST, hand) + CONTACT_ENERGY PRESET."""
    e, o = ctx["eng"], ctx["osc"]
    legal_o, energy_o, dist_o, order, tf7 = o.eval(0, ctx["obj"], ctx["raw"], mode=0, want_dist=True)
    res = e.eval_batch(ctx["rid"], ctx["sb"].body_ids[ctx["obj"]], ctx["raw"], pair_mask=True, link_dist=True, body_tf=True)
    # FK (A11): fp64 on both sides
    assert np.abs(res["body_tf"][:, :, 4:7] - tf7[:, :, 4:7]).max() < 1e-9
    dq = np.minimum(np.abs(res["body_tf"][:, :, 0:4] - tf7[:, :, 0:4]).max(-1), np.abs(res["body_tf"][:, :, 0:4] + tf7[:, :, 0:4]).max(-1))
    assert dq.max() < 1e-12
    # legality (A5/A6/A13): exact
    assert (res["legal"] == legal_o).all()
    assert 100 < legal_o.sum() < 900
    # energy (A8/A13)
    m = legal_o == 1
    assert relerr(res["energy"][m], energy_o[m]).max() < REL
    assert (res["energy"][~m] == 0).all()
    # link distances (A7): -1 on interpenetration, else within tolerance
    ld = res["link_dist"].astype(np.float64)
    assert ((ld < 0) == (dist_o < 0)).all()
    both = dist_o >= 0
    assert relerr(ld[both], dist_o[both]).max() < REL
    ctx["res"] = res


def test_config1_collision_pair_sets(ctx):
    """allCollisions(ALL_COLLISIONS, interest = hand): pair sets equal as sets (reference order is pointer order)."""
    e, o, sb = ctx["eng"], ctx["osc"], ctx["sb"]
    n = 400
    raw = ctx["raw"][:n]
    res = e.eval_batch(ctx["rid"], sb.body_ids[ctx["obj"]], raw, pair_mask=True)
    pairs = e.robot_pairs(ctx["rid"])
    inv = {v: k for k, v in sb.body_ids.items()}
    want = o.pair_sets(0, raw)
    ncol = 0
    for p in range(n):
        got = {tuple(sorted((inv[a], inv[b]))) for i, (a, b) in enumerate(pairs)
               if (int(res["pair_mask"][p, i >> 6]) >> (i & 63)) & 1}
        assert got == want[p], (p, got ^ want[p])
        ncol += len(got)
        assert (len(got) == 0) == bool(res["legal"][p])
    assert ncol > 100


def test_config1_energy_live(ctx):
    """CONTACT_LIVE: virtual contacts rebuilt from bodyToBodyDistance per link, incl. the reference's frame quirks."""
    e, o = ctx["eng"], ctx["osc"]
    legal_o, energy_o, _, _, _ = o.eval(0, ctx["obj"], ctx["raw"], mode=1)
    res = e.eval_batch(ctx["rid"], ctx["sb"].body_ids[ctx["obj"]], ctx["raw"], live=True)
    assert (res["legal"] == legal_o).all()
    m = legal_o == 1
    assert relerr(res["energy"][m], energy_o[m]).max() < REL


def test_search_state_input_matches_raw_input(ctx):
    """B2C_INPUT_AA_EIGEN (A12): SPACE_AXIS_ANGLE + POSE_EIGEN variables mapped on the device."""
    from oracle.port import fk as ofk
    sc, e, o = ctx["sc"], ctx["eng"], ctx["osc"]
    r = sc.robots[0]
    rng = np.random.default_rng(5)
    n = 600
    amps = rng.uniform(-4, 4, size=(n, r.eg.shape[0])).astype(np.float32)
    states = np.concatenate([ctx["pos"][:n], amps], 1).astype(np.float32)
    s64 = states.astype(np.float64)
    base = ofk.aa_state_to_base(r, sc.body_tran[ctx["obj"]], s64[:, :6])
    dofs = ofk.eigen_to_dof(r, s64[:, 6:])
    raw64 = np.concatenate([base[0], base[1], dofs], 1)
    legal_o, energy_o, _, order, tf7 = o.eval(0, ctx["obj"], raw64, mode=0)
    res = e.eval_batch(ctx["rid"], ctx["sb"].body_ids[ctx["obj"]], states, input_mode=B2C_INPUT_AA_EIGEN,
                       ref_tran=sc.body_tran[ctx["obj"]], body_tf=True)
    assert np.abs(res["body_tf"][:, :, 4:7] - tf7[:, :, 4:7]).max() < 1e-8
    assert (res["legal"] == legal_o).all()
    m = legal_o == 1
    assert relerr(res["energy"][m], energy_o[m]).max() < REL


@pytest.mark.parametrize("kind", ["ellipsoid", "ellipsoid_abc", "approach", "complete_dof"])
def test_other_search_state_layouts(ctx, kind):
    """A12: SPACE_ELLIPSOID / SPACE_APPROACH + POSE_EIGEN and SPACE_COMPLETE + POSE_DOF mapped on the device; the oracle is
    the numpy restatement, itself pinned to the reference's own getCoreTran lines (tests/test_oracle_fk_pinned.py)."""
    from graspit_b200.engine import B2C_INPUT_APPROACH_EIGEN, B2C_INPUT_COMPLETE_DOF, B2C_INPUT_ELLIPSOID_EIGEN
    from oracle.port import fk as ofk
    sc, e, o = ctx["sc"], ctx["eng"], ctx["osc"]
    r = sc.robots[0]
    ref_tran = sc.body_tran[ctx["obj"]]
    rng = np.random.default_rng(17)
    n = 400
    amps = rng.uniform(-4, 4, size=(n, r.eg.shape[0]))
    params = None
    if kind.startswith("ellipsoid"):
        pos = np.stack([rng.uniform(-np.pi / 2, np.pi / 2, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(-50, 100, n)], 1)
        params = (60.0, 90.0, 120.0) if kind == "ellipsoid_abc" else None
        states = np.concatenate([pos, amps], 1).astype(np.float32)
        s64 = states.astype(np.float64)
        base = ofk.ellipsoid_state_to_base(r, ref_tran, s64[:, :4], abc=params or (80.0, 80.0, 160.0))
        dofs = ofk.eigen_to_dof(r, s64[:, 4:]); mode = B2C_INPUT_ELLIPSOID_EIGEN
    elif kind == "approach":
        pos = np.stack([rng.uniform(-30, 200, n), rng.uniform(-np.pi / 3, np.pi / 3, n), rng.uniform(-np.pi / 3, np.pi / 3, n)], 1)
        pos[0, 1] = 0.0; pos[1, 2] = 2.0e-4                       # below the AXIS_ANGLE_ROTATION identity snap
        states = np.concatenate([pos, amps], 1).astype(np.float32)
        s64 = states.astype(np.float64)
        base = ofk.approach_state_to_base(r, ref_tran, s64[:, :3])
        dofs = ofk.eigen_to_dof(r, s64[:, 3:]); mode = B2C_INPUT_APPROACH_EIGEN
    else:
        q = rng.uniform(-5, 5, size=(n, 4))
        pos = np.concatenate([rng.uniform(-250, 250, (n, 2)), rng.uniform(-150, 350, (n, 1)), q], 1)
        d = rng.uniform(r.dof_min, r.dof_max, size=(n, r.n_dof))
        states = np.concatenate([pos, d], 1).astype(np.float32)
        s64 = states.astype(np.float64)
        base = ofk.complete_state_to_base(ref_tran, s64[:, :7])
        dofs = s64[:, 7:]; mode = B2C_INPUT_COMPLETE_DOF
    raw64 = np.concatenate([base[0], base[1], dofs], 1)
    legal_o, energy_o, _, order, tf7 = o.eval(0, ctx["obj"], raw64, mode=0)
    res = e.eval_batch(ctx["rid"], ctx["sb"].body_ids[ctx["obj"]], states, input_mode=mode, ref_tran=ref_tran, body_tf=True,
                       position_params=params)
    assert np.abs(res["body_tf"][:, :, 4:7] - tf7[:, :, 4:7]).max() < 1e-8
    dq = np.minimum(np.abs(res["body_tf"][:, :, 0:4] - tf7[:, :, 0:4]).max(-1), np.abs(res["body_tf"][:, :, 0:4] + tf7[:, :, 0:4]).max(-1))
    assert dq.max() < 1e-11
    assert (res["legal"] == legal_o).all()
    m = legal_o == 1
    assert m.sum() > 10
    assert relerr(res["energy"][m], energy_o[m]).max() < REL


def test_reference_collision_tests_through_the_abi(ctx):
    """test/graspit_collision_tests.cpp, all three tests, against the CUDA backend."""
    sc = ctx["sc"]
    eng = Engine(0)
    mug = eng.add_body(sc.bodies[0].tris)
    palm = eng.add_body(sc.bodies[1].tris)
    rot = ref.transf_axis_angle(10.0, [1, 0, 0])
    tr = ref.transf_compose(([1, 0, 0, 0], [1, 2, 3]), rot)
    eng.set_body_transform(mug, tr[0], tr[1])
    d, p1, p2 = eng.body_to_body_distance(mug, palm)
    assert d == pytest.approx(-1, abs=1e-3)
    assert p1 == pytest.approx([-1, 3.31021, 1.42917], abs=1e-3)
    assert p2 == pytest.approx([0, 0, 0], abs=1e-3)
    d, cp, cn = eng.point_to_body_distance(mug, [10, 11, 12])
    assert d == pytest.approx(33.8332, abs=1e-3)
    assert cn == pytest.approx([0.951057, -0.168109, 0.259287], abs=1e-3)
    assert cp == pytest.approx([42.1773, 5.31235, 20.7725], abs=1e-3)
    bm = eng.get_bounding_volumes(mug, 0)
    assert bm.shape == (1, 10)
    assert bm[0, 0:3] == pytest.approx([-11.662, -8.98983, 0.000190189], abs=1e-3)
    assert bm[0, 3:7] == pytest.approx([0.999999, -9.15765e-06, -1.47701e-06, 0.00157698], abs=1e-3)
    bp = eng.get_bounding_volumes(palm, 0)
    assert bp[0, 0:3] == pytest.approx([0.013166, -9.91786, -38.5818], abs=1e-3)
    assert bp[0, 3:7] == pytest.approx([0.6664499, 0.236398, -0.66585779, -0.23789390], abs=1e-3)
    eng.close()


def test_single_pose_queries_match_oracle(ctx):
    """bodyToBodyDistance (incl. the double-inverse output quirk), pointToBodyDistance, allCollisions."""
    sc = ctx["sc"]
    eng = Engine(0)
    w = ref.RefWorld()
    ids_e = [eng.add_body(b.tris) for b in sc.bodies[:3]]
    ids_o = [w.add_body(b.tris.astype(np.float64)) for b in sc.bodies[:3]]
    rng = np.random.default_rng(11)
    nhit = 0
    for trial in range(40):
        for k in range(3):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            t = rng.uniform(-60, 60, 3) if trial % 2 else rng.uniform(-140, 140, 3)
            eng.set_body_transform(ids_e[k], q, t); w.set_transform(ids_o[k], q, t)
        n_o, pairs_o = w.all_collisions(fast=False)
        n_e, pairs_e = eng.all_collisions(ALL_COLLISIONS)
        inv_e = {v: i for i, v in enumerate(ids_e)}; inv_o = {v: i for i, v in enumerate(ids_o)}
        assert {tuple(sorted((inv_e[a], inv_e[b]))) for a, b in pairs_e} == {tuple(sorted((inv_o[a], inv_o[b]))) for a, b in pairs_o}
        assert n_e == n_o
        assert eng.all_collisions(FAST_COLLISION)[0] == (1 if n_o else 0)
        nhit += n_o
        d_o, p1_o, p2_o = w.body_distance(ids_o[1], ids_o[0])
        d_e, p1_e, p2_e = eng.body_to_body_distance(ids_e[1], ids_e[0])
        if d_o < 0:
            assert d_e == -1
            assert np.allclose(p1_e, p1_o, atol=1e-9) and np.allclose(p2_e, p2_o, atol=1e-9)
        else:
            assert abs(d_e - d_o) <= REL * d_o
            # closest points may legitimately differ on ties; their separation must equal the distance
        p = rng.uniform(-200, 200, 3)
        dd_o, cp_o, cn_o = w.point_distance(ids_o[0], p)
        dd_e, cp_e, cn_e = eng.point_to_body_distance(ids_e[0], p)
        assert abs(dd_e - dd_o) <= 1e-6 * max(dd_o, 1e-9)
        assert np.allclose(cp_e, cp_o, atol=1e-5) and np.allclose(cn_e, cn_o, atol=1e-6)
    assert nhit > 3
    eng.close()


def test_activation_interest_and_clones(ctx):
    sc = ctx["sc"]
    eng = Engine(0)
    a = eng.add_body(sc.bodies[1].tris)
    b = eng.clone_body(a)                      # cloneBody: shares geometry, own transform
    c = eng.add_body(sc.bodies[0].tris)
    for x in (a, b):
        eng.set_body_transform(x, [1, 0, 0, 0], [0, 0, 0])
    eng.set_body_transform(c, [1, 0, 0, 0], [0, 0, 500])
    assert eng.all_collisions(ALL_COLLISIONS)[1] == [(a, b)]          # identical bodies on top of each other
    assert eng.is_active(a, b) and eng.is_active(a)
    eng.activate_pair(a, b, False)
    assert not eng.is_active(a, b)
    assert eng.all_collisions(ALL_COLLISIONS)[0] == 0
    eng.activate_pair(a, b, True)
    eng.activate_body(b, False)
    assert not eng.is_active(b) and not eng.is_active(a, b)
    assert eng.all_collisions(FAST_COLLISION)[0] == 0
    eng.activate_body(b, True)
    assert eng.all_collisions(ALL_COLLISIONS, interest=[c])[0] == 0   # interest list: only pairs touching c
    assert eng.all_collisions(ALL_COLLISIONS, interest=[a])[0] == 1
    eng.remove_body(b)
    assert eng.all_collisions(FAST_COLLISION)[0] == 0
    # unknown bodies: reference logs "model not found" and returns 0 / false
    assert eng.body_to_body_distance(a, 99)[0] == 0.0
    assert eng.point_to_body_distance(99, [0, 0, 0])[0] == 0.0
    assert not eng.is_active(99)
    eng.close()


def test_edge_cases(ctx):
    sc = ctx["sc"]
    eng = Engine(0)
    one = eng.add_body(np.array([[[0, 0, 0], [10, 0, 0], [0, 10, 0]]], dtype=np.float32))        # single triangle
    two = eng.add_body(np.array([[[1, 1, -5], [1, 1, 5], [8, 1, 5]]], dtype=np.float32))         # pierces it
    empty = eng.add_body(np.zeros((0, 3, 3), dtype=np.float32))                                  # no geometry
    n, pairs = eng.all_collisions(ALL_COLLISIONS)
    assert pairs == [(one, two)]
    eng.set_body_transform(two, [1, 0, 0, 0], [0, 0, 20])
    assert eng.all_collisions(ALL_COLLISIONS)[0] == 0
    d, _, _ = eng.body_to_body_distance(one, two, local=True)
    assert d == pytest.approx(15.0, rel=1e-6)
    # touching counts as intersecting (reference project6 uses strict >)
    eng.set_body_transform(two, [1, 0, 0, 0], [0, 0, 5])
    assert eng.all_collisions(ALL_COLLISIONS)[0] == 1
    assert eng.body_to_body_distance(one, two)[0] == -1
    # empty batch
    sb = ctx["sb"]
    res = ctx["eng"].eval_batch(ctx["rid"], sb.body_ids[ctx["obj"]], np.zeros((0, 11), dtype=np.float32))
    assert res["legal"].shape == (0,)
    with pytest.raises(B2CError):
        ctx["eng"].eval_batch(ctx["rid"], 12345, ctx["raw"][:4])
    with pytest.raises(B2CError):
        ctx["eng"].eval_batch(ctx["rid"], sb.body_ids[ctx["obj"]], ctx["raw"][:4, :9])      # stride too small
    eng.close()


def test_full_size_properties(ctx):
    """BASELINE config 2 size (65 536 states): determinism, permutation equivariance, rigid-motion invariance."""
    sc, e, sb = ctx["sc"], ctx["eng"], ctx["sb"]
    n = 65536
    pos, dofs = sample_aa_states(sc, n, seed=4321)
    raw = aa_to_raw_states(sc, pos, dofs)
    oid = sb.body_ids[ctx["obj"]]
    r1 = e.eval_batch(ctx["rid"], oid, raw)
    r2 = e.eval_batch(ctx["rid"], oid, raw)
    assert (r1["legal"] == r2["legal"]).all() and np.array_equal(r1["energy"], r2["energy"])
    perm = np.random.default_rng(0).permutation(n)
    r3 = e.eval_batch(ctx["rid"], oid, raw[perm])
    assert (r3["legal"] == r1["legal"][perm]).all() and np.array_equal(r3["energy"], r1["energy"][perm])
    assert (r1["energy"][r1["legal"] == 0] == 0).all() and (r1["energy"][r1["legal"] == 1] > 0).all()
    # spot-check 256 of them against the oracle
    idx = np.random.default_rng(1).choice(n, 256, replace=False)
    legal_o, energy_o, _, _, _ = ctx["osc"].eval(0, ctx["obj"], raw[idx], mode=0)
    assert (r1["legal"][idx] == legal_o).all()
    m = legal_o == 1
    assert relerr(r1["energy"][idx][m], energy_o[m]).max() < REL
    # move the whole world rigidly: results must not change (beyond fp32 tolerance of the energy)
    from oracle.port import transf_np as T
    g = T.make(np.array([[0.3, -0.5, 0.7, 0.4]]), np.array([[40.0, -25.0, 60.0]]))
    k = 4096
    base = T.make(raw[:k, 0:4].astype(np.float64), raw[:k, 4:7].astype(np.float64))
    moved = T.compose((np.repeat(g[0], k, 0), np.repeat(g[1], k, 0)), base)
    raw_m = np.concatenate([moved[0], moved[1], raw[:k, 7:]], 1).astype(np.float32)
    oq, ot = e.get_body_transform(oid)
    om = T.compose(g, T.make(oq[None], ot[None]))
    e.set_body_transform(oid, om[0][0], om[1][0])
    try:
        r4 = e.eval_batch(ctx["rid"], oid, raw_m)
    finally:
        e.set_body_transform(oid, oq, ot)
    # fp32 rounding of the moved base can flip poses that are within ~1e-4 mm of tangency
    assert (r4["legal"] != r1["legal"][:k]).mean() < 2e-3
    both = (r4["legal"] == 1) & (r1["legal"][:k] == 1)
    assert np.median(relerr(r4["energy"][both], r1["energy"][:k][both])) < 1e-5


def test_small_distances_many_poses(ctx):
    """Regression: many close poses (distances from 1e-3 mm to a few mm), where the fp64 refinement of tiny
    distances and the fp32 traversal must agree on which triangle pair wins."""
    sc, e, o = ctx["sc"], ctx["eng"], ctx["osc"]
    pos, dofs = sample_aa_states(sc, 3000, seed=99, strata=("close",))
    raw = aa_to_raw_states(sc, pos, dofs)
    _, _, dist_o, _, _ = o.eval(0, ctx["obj"], raw, mode=0, want_dist=True)
    res = e.eval_batch(ctx["rid"], ctx["sb"].body_ids[ctx["obj"]], raw, link_dist=True)
    ld = res["link_dist"].astype(np.float64)
    assert ((ld < 0) == (dist_o < 0)).all()
    both = dist_o >= 0
    assert (dist_o[both] < 1.0).sum() > 50
    assert relerr(ld[both], dist_o[both]).max() < REL


def test_large_random_batches(ctx):
    """20 000 poses (PRESET) and 4 000 poses (LIVE) in the planner's state layout against the reference: legality exact,
    energies within the stated tolerance -- a wider net than config 1 for rare traversal / precision paths."""
    from oracle.port import fk as ofk
    sc, e, o = ctx["sc"], ctx["eng"], ctx["osc"]
    r = sc.robots[0]
    rng = np.random.default_rng(77)
    n = 20000
    pos, _ = sample_aa_states(sc, n, seed=4242, strata=("far", "near", "close"))
    states = np.concatenate([pos, rng.uniform(-4, 4, size=(n, r.eg.shape[0]))], 1).astype(np.float32)
    s64 = states.astype(np.float64)
    base = ofk.aa_state_to_base(r, sc.body_tran[ctx["obj"]], s64[:, :6])
    raw64 = np.concatenate([base[0], base[1], ofk.eigen_to_dof(r, s64[:, 6:])], 1)
    legal_o, energy_o, _, _, _ = o.eval(0, ctx["obj"], raw64, mode=0)
    res = e.eval_batch(ctx["rid"], ctx["sb"].body_ids[ctx["obj"]], states, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[ctx["obj"]])
    assert (res["legal"] == legal_o).all()
    m = legal_o == 1
    assert 3000 < m.sum() < 19000
    assert relerr(res["energy"][m], energy_o[m]).max() < REL
    k = 4000
    legal_l, energy_l, _, _, _ = o.eval(0, ctx["obj"], raw64[:k], mode=1)
    live = e.eval_batch(ctx["rid"], ctx["sb"].body_ids[ctx["obj"]], states[:k], input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[ctx["obj"]], live=True)
    assert (live["legal"] == legal_l).all()
    ml = legal_l == 1
    assert relerr(live["energy"][ml], energy_l[ml]).max() < REL
