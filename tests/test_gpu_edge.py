"""Edge cases and full-size properties of the batched path: empty / single / ragged inputs, error behaviour, the
ListPlanner file workflow, and BASELINE.json's largest configuration (1 048 576 poses vs the 220 800-triangle mug)
checked through size-independent properties."""
import io

import numpy as np
import pytest

from graspit_b200.engine import B2C_INPUT_AA_EIGEN, B2CError, Engine, SceneBinding
from graspit_b200.formats import statefile as sf
from graspit_b200.formats.scenepack import load_pack
from graspit_b200.planner import BEST_LIST_SIZE, ListPlanner
from oracle import ref
from tests.common import OracleScene, object_body, sample_aa_states

pytestmark = pytest.mark.gpu
REL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    obj = object_body(sc)
    return dict(sc=sc, eng=eng, sb=sb, obj=obj, rid=sb.robot_ids[0], oid=sb.body_ids[obj], ref=sc.body_tran[obj])


def _states(sc, n, seed):
    pos, _ = sample_aa_states(sc, n, seed=seed)
    amps = np.random.default_rng(seed + 1).uniform(-4, 4, size=(n, sc.robots[0].eg.shape[0]))
    return np.concatenate([pos, amps], 1).astype(np.float32)


def test_empty_and_single_pose_batches(ctx):
    eng, st = ctx["eng"], _states(ctx["sc"], 5, 3)
    r0 = eng.eval_batch(ctx["rid"], ctx["oid"], np.zeros((0, 8), np.float32), input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"])
    assert len(r0["legal"]) == 0 and len(r0["energy"]) == 0
    r5 = eng.eval_batch(ctx["rid"], ctx["oid"], st, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"], live=True, link_dist=True, pair_mask=True)
    for i in range(5):                                 # a batch of one equals row i of the batch of five
        r1 = eng.eval_batch(ctx["rid"], ctx["oid"], st[i:i + 1], input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"], live=True, link_dist=True, pair_mask=True)
        assert r1["legal"][0] == r5["legal"][i] and r1["energy"][0] == r5["energy"][i]
        assert (r1["link_dist"][0] == r5["link_dist"][i]).all() and (r1["pair_mask"][0] == r5["pair_mask"][i]).all()


def test_argument_errors_are_reported_not_crashed(ctx):
    eng, st = ctx["eng"], _states(ctx["sc"], 4, 3)
    with pytest.raises(B2CError):
        eng.eval_batch(99, ctx["oid"], st, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"])          # unknown robot
    with pytest.raises(B2CError):
        eng.eval_batch(ctx["rid"], 9999, st, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"])        # unknown object
    with pytest.raises(B2CError):
        eng.eval_batch(ctx["rid"], ctx["oid"], st[:, :5], input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"])   # stride too small
    with pytest.raises(B2CError):
        eng.eval_batch(ctx["rid"], -1, st, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"], live=True)       # LIVE needs an object
    # reference behaviour for unknown bodies: 0 / false, no exception through the CollisionInterface mirror
    assert eng.body_to_body_distance(12345, ctx["oid"])[0] == 0
    assert eng.is_active(12345) is False
    # the context is still usable afterwards
    assert len(eng.eval_batch(ctx["rid"], ctx["oid"], st, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"])["legal"]) == 4


def test_empty_mesh_body_never_collides_and_is_far_away():
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    mug = eng.add_body(sc.bodies[0].tris)
    empty = eng.add_body(np.zeros((0, 3, 3), np.float32))
    n, pairs = eng.all_collisions()
    assert n == 0 and pairs == []
    d = eng.body_to_body_distance(mug, empty)[0]
    assert d > 1.0e15
    assert eng.all_contacts(0.2) == []
    assert len(eng.body_region(empty, [0, 0, 0], [0, 0, 1], 5.0)) == 1      # only the reference point itself
    # interest list that names no known body: nothing is tested
    assert eng.all_collisions(interest=[]) == (0, [])


def test_list_planner_from_a_state_file(ctx, tmp_path):
    """ListPlanner workflow: states -> text file (reference wire format) -> batch evaluation -> best list."""
    sc, eng = ctx["sc"], ctx["eng"]
    st = _states(sc, 3000, 17)
    path = str(tmp_path / "states.txt")
    sf.write_states(path, [sf.PlannerState(sf.POSE_EIGEN, r[6:], sf.SPACE_AXIS_ANGLE, r[:6]) for r in st])
    states = sf.read_states(path, n_posture=2)
    assert len(states) == 3000
    lp = ListPlanner(eng, ctx["rid"], ctx["oid"], ctx["ref"])
    best = lp.run(states)
    # the file carries six decimals: evaluate exactly what was read
    _, arr = sf.to_engine_batch(states)
    direct = eng.eval_batch(ctx["rid"], ctx["oid"], arr, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ctx["ref"])
    assert (lp.legal == direct["legal"]).all() and (lp.energy == direct["energy"]).all()
    legal_idx = np.nonzero(direct["legal"])[0]
    order = legal_idx[np.argsort(direct["energy"][legal_idx], kind="stable")][:BEST_LIST_SIZE]
    assert [i for _, i in best] == [int(i) for i in order]
    assert len(best) == BEST_LIST_SIZE and best[0][0] <= best[-1][0] < 1.0e5
    # and against the reference on a sample
    from oracle.port import fk as ofk
    r = sc.robots[0]
    s64 = arr[:300].astype(np.float64)
    base = ofk.aa_state_to_base(r, sc.body_tran[ctx["obj"]], s64[:, :6])
    raw64 = np.concatenate([base[0], base[1], ofk.eigen_to_dof(r, s64[:, 6:])], 1)
    osc = OracleScene(sc, nthreads=min(8, ref.hardware_threads()))
    legal_o, energy_o, _, _, _ = osc.eval(0, ctx["obj"], raw64, mode=0)
    assert (lp.legal[:300] == legal_o).all()
    m = legal_o == 1
    assert (np.abs(lp.energy[:300][m] - energy_o[m]) <= REL * np.abs(energy_o[m])).all()


def test_config5_full_size_properties():
    """1 048 576 poses vs the 220 800-triangle mug in one call (BASELINE.json configs[4] per-GPU upper size):
    deterministic (idempotent), independent of how the batch is split (what sharding over GPUs relies on), and equal
    to the 3 450-triangle mug on a sample (subdivision does not move the surface)."""
    big, small = load_pack("barrett_mug_200k"), load_pack("barrett_mug")
    eng, eng_s = Engine(0), Engine(0)
    sb, sb_s = SceneBinding(eng, big), SceneBinding(eng_s, small)
    obj = object_body(big)
    n = 1 << 20
    st = _states(big, n, 2024)
    args = dict(input_mode=B2C_INPUT_AA_EIGEN, ref_tran=big.body_tran[obj])
    whole = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st, **args)
    assert 0.2 < whole["legal"].mean() < 0.9
    again = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st, **args)
    assert (again["legal"] == whole["legal"]).all() and (again["energy"] == whole["energy"]).all()
    for k in range(8):                                                     # 8 shards, as on 8 GPUs
        lo, hi = k * n // 8, (k + 1) * n // 8
        part = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st[lo:hi], **args)
        assert (part["legal"] == whole["legal"][lo:hi]).all() and (part["energy"] == whole["energy"][lo:hi]).all()
    idx = np.random.default_rng(5).choice(n, 30000, replace=False)
    ref_small = eng_s.eval_batch(sb_s.robot_ids[0], sb_s.body_ids[obj], st[idx], **args)
    flips = (ref_small["legal"] != whole["legal"][idx]).sum()
    assert flips <= 3                                                      # fp32 midpoints: tangency within an ulp
    m = (ref_small["legal"] == 1) & (whole["legal"][idx] == 1)
    e1, e2 = ref_small["energy"][m].astype(np.float64), whole["energy"][idx][m].astype(np.float64)
    assert (np.abs(e1 - e2) <= REL * np.abs(e1) + 1e-4).all()
    assert (whole["energy"][whole["legal"] == 0] == 0).all()


def test_batched_grasp_entry_points_edge_cases(ctx):
    """autograsp / contacts / region batches: empty inputs, argument errors, a batch of one equals its row of a larger batch"""
    from graspit_b200.engine import AG_INTERP_ERROR, B2C_INPUT_JOINTS
    from tests.common import aa_to_raw_states
    eng, sc, rid = ctx["eng"], ctx["sc"], ctx["rid"]
    r = sc.robots[0]
    pos, dofs = sample_aa_states(sc, 64, seed=9, strata=("near",))
    dofs = dofs.copy(); dofs[:, 1:] = 0.0
    raw = aa_to_raw_states(sc, pos, dofs)
    # empty batches
    out0 = eng.auto_grasp(rid, r, raw[:0])
    assert out0["dof"].shape == (0, r.n_dof) and out0["rounds"] == 0
    c0 = eng.contacts_batch(rid, raw[:0], 0.2)
    assert len(c0["contacts"]) == 0 and list(c0["offsets"]) == [0]
    assert eng.body_region_batch([], np.zeros((0, 3)), np.zeros((0, 3)), []) == []
    # errors are reported, the context survives
    with pytest.raises(B2CError):
        eng.auto_grasp(99, r, raw)
    with pytest.raises(B2CError):
        eng.contacts_batch(rid, raw[:, :8], 0.2)                                   # stride too small
    with pytest.raises(B2CError):
        eng.contacts_batch(rid, np.zeros((4, 9)), 0.2, input_mode=B2C_INPUT_JOINTS)  # 7 + 8 joints needed
    with pytest.raises(B2CError):
        eng.body_region_batch([12345], np.zeros((1, 3)), np.array([[0, 0, 1.0]]), [3.0])
    # a batch of one = its row of the batch of 64 (poses never interact)
    full = eng.auto_grasp(rid, r, raw)
    for i in (0, 17, 63):
        one = eng.auto_grasp(rid, r, raw[i:i + 1])
        assert np.array_equal(one["dof"][0], full["dof"][i]) and one["status"][0] == full["status"][i]
        assert np.array_equal(one["joints"][0], full["joints"][i]) and one["iters"][0] == full["iters"][i]
    ok = (full["status"] & AG_INTERP_ERROR) == 0
    js = np.concatenate([raw[ok, :7].astype(np.float64), full["joints"][ok]], axis=1)
    call = eng.contacts_batch(rid, js, 0.2, input_mode=B2C_INPUT_JOINTS)
    for i in (0, 5, len(js) - 1):
        one = eng.contacts_batch(rid, js[i:i + 1], 0.2, input_mode=B2C_INPUT_JOINTS)
        lo, hi = call["offsets"][i], call["offsets"][i + 1]
        assert np.array_equal(one["contacts"], call["contacts"][lo:hi]) and np.array_equal(one["pair"], call["pair"][lo:hi])
    # truncation: a small cap still reports the full count through offsets (the wrapper retries with the right size)
    small = eng.contacts_batch(rid, js, 0.2, input_mode=B2C_INPUT_JOINTS, cap_contacts=3)
    assert np.array_equal(small["contacts"], call["contacts"])


def test_two_devices_in_one_process():
    """one process may hold a context per GPU (dynamic shared-memory opt-in is per device): same results on both"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = load_pack("barrett_mug")
    obj = object_body(sc)
    st = _states(sc, 512, 11)
    out = []
    for dev in (1, 0):                       # the second device first: nothing may depend on device 0 having run
        eng = Engine(dev)
        sb = SceneBinding(eng, sc)
        out.append(eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[obj],
                                  live=True, link_dist=True))
    assert (out[0]["legal"] == out[1]["legal"]).all() and np.array_equal(out[0]["energy"], out[1]["energy"])
    assert np.array_equal(out[0]["link_dist"], out[1]["link_dist"])
