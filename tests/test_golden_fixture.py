"""Committed golden vectors (tests/golden/, made by tools/make_golden.py from the reference's own GraspitCollision):
the oracle must still reproduce them (CPU), and the CUDA path must match them through the C ABI (GPU)."""
import os

import numpy as np
import pytest

from graspit_b200.formats.scenepack import load_pack

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "barrett_mug_config1.npz")
REL = 1e-4


def _fixture():
    z = np.load(FIX)
    shape = tuple(int(x) for x in z["pair_shape"])
    pm = np.unpackbits(z["pair_bitmap"])[: int(np.prod(shape))].reshape(shape)
    return z, pm


def test_oracle_reproduces_the_golden_vectors():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    from tests.common import OracleScene, object_body
    z, pm = _fixture()
    sc = load_pack("barrett_mug")
    osc = OracleScene(sc, nthreads=min(8, ref.hardware_threads()))
    n = 250
    legal, energy, dist, order, tf7 = osc.eval(0, object_body(sc), z["raw"][:n], mode=0, want_dist=True)
    assert (legal == z["legal"][:n]).all()
    assert np.array_equal(energy, z["energy_preset"][:n]) and np.array_equal(dist, z["link_dist"][:n])
    assert list(order) == list(z["body_order"]) and np.array_equal(tf7, z["body_tf"][:n])
    pairs = osc.pair_sets(0, z["raw"][:40])
    for p, s in enumerate(pairs):
        assert {(a, b) for a, b in zip(*np.nonzero(pm[p]))} == s


@pytest.mark.gpu
def test_cuda_path_matches_the_golden_vectors():
    from graspit_b200.engine import Engine, SceneBinding
    from tests.common import object_body
    z, pm = _fixture()
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    obj = object_body(sc)
    res = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], z["raw"], pair_mask=True, link_dist=True, body_tf=True)
    live = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], z["raw"], live=True)
    assert (res["legal"] == z["legal"]).all() and (live["legal"] == z["legal"]).all()
    m = z["legal"] == 1
    rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-9)
    assert rel(res["energy"][m], z["energy_preset"][m]).max() < REL          # tolerance: BASELINE.json north_star
    assert rel(live["energy"][m], z["energy_live"][m]).max() < REL
    ld, want = res["link_dist"].astype(np.float64), z["link_dist"]
    assert ((ld < 0) == (want < 0)).all()
    both = want >= 0
    assert rel(ld[both], want[both]).max() < REL
    assert np.abs(res["body_tf"][:, :, 4:7] - z["body_tf"][:, :, 4:7]).max() < 1e-9
    pairs = eng.robot_pairs(sb.robot_ids[0])
    inv = {v: k for k, v in sb.body_ids.items()}
    for p in range(len(z["raw"])):
        got = {tuple(sorted((inv[a], inv[b]))) for i, (a, b) in enumerate(pairs) if (int(res["pair_mask"][p, i >> 6]) >> (i & 63)) & 1}
        assert got == {(int(a), int(b)) for a, b in zip(*np.nonzero(pm[p]))}, p


AG_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "barrett_mug_autograsp.npz")


def test_autograsp_oracle_reproduces_the_golden_vectors():
    """tools/make_golden.py autograsp: the restated control flow on the reference backend is deterministic"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    from oracle.port.autograsp import AutoGraspOracle
    from tests.common import OracleScene
    z = np.load(AG_FIX)
    sc = load_pack("barrett_mug")
    osc = OracleScene(sc, nthreads=1)
    ag = AutoGraspOracle(sc, 0, osc.w, osc.ids)
    for k in (0, 57, 124):
        raw = z["raw"][k]
        r = ag.auto_grasp((raw[0:4].astype(np.float64), raw[4:7].astype(np.float64)), raw[7:].astype(np.float64))
        assert r["error"] == bool(z["error"][k]) and r["moved"] == bool(z["moved"][k]) and r["iters"] == int(z["iters"][k])
        if not r["error"]:
            assert np.array_equal(r["dof"], z["dof"][k]) and np.array_equal(r["stopped"], z["stopped"][k])


@pytest.mark.gpu
def test_cuda_autograsp_matches_the_golden_vectors():
    from graspit_b200.engine import AG_INTERP_ERROR, AG_MOVED, Engine, SceneBinding
    z = np.load(AG_FIX)
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    out = eng.auto_grasp(sb.robot_ids[0], sc.robots[0], z["raw"])
    err = (out["status"] & AG_INTERP_ERROR) != 0
    assert np.array_equal(err, z["error"]) and np.array_equal((out["status"] & AG_MOVED) != 0, z["moved"])
    ok = ~z["error"]
    d = np.abs(out["dof"][ok] - z["dof"][ok]).max(axis=1)
    # identical decisions => identical bisection parameters; a decision inside fp32 rounding of a threshold moves one level
    assert (d < 1e-9).mean() >= 0.9 and (d < 2e-3).mean() >= 0.97 and d.max() < 2e-2
    same = d < 1e-9
    assert np.array_equal(out["stopped"][ok][same], z["stopped"][ok][same])
    assert np.array_equal(out["iters"][ok][same], z["iters"][ok][same])
    assert np.abs(out["joints"][ok][same] - z["joints"][ok][same]).max() < 1e-9
