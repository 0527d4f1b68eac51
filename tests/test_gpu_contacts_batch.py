"""GPU: batched contacts (b2c_contacts_batch = allContacts(threshold, interest = hand) per pose) and the joint-level
input mode, on hands closed by the batched autograsp: against the reference backend, and against the single-pose path."""
import os

import numpy as np
import pytest

from graspit_b200.engine import AG_INTERP_ERROR, B2C_INPUT_JOINTS, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle.port.autograsp import AutoGraspOracle
from tests.common import OracleScene, object_body

pytestmark = pytest.mark.gpu
AG_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "barrett_mug_autograsp.npz")
THR = 0.2


@pytest.fixture(scope="module")
def closed():
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    rid = sb.robot_ids[0]
    raw = np.load(AG_FIX)["raw"]
    out = eng.auto_grasp(rid, sc.robots[0], raw)
    ok = (out["status"] & AG_INTERP_ERROR) == 0
    js = np.concatenate([raw[ok, :7].astype(np.float64), out["joints"][ok]], axis=1)
    return dict(sc=sc, eng=eng, sb=sb, rid=rid, js=js, raw=raw[ok], dof=out["dof"][ok])


def _by_pose_pair(res, pairs, inv):
    out = {}
    for c, p, q in zip(res["contacts"], res["pose"], res["pair"]):
        a, b = pairs[q]
        out.setdefault((int(p), tuple(sorted((inv[a], inv[b])))), []).append((inv[a] > inv[b], c))
    return out


def test_joint_input_fk_and_legality(closed):
    sc, eng, sb, rid, js = closed["sc"], closed["eng"], closed["sb"], closed["rid"], closed["js"]
    obj = sb.body_ids[object_body(sc)]
    res = eng.eval_batch(rid, obj, js, input_mode=B2C_INPUT_JOINTS, body_tf=True, link_dist=True)
    # the bisection leaves every finger 0 < d < THRESHOLD / 2 short of touching: collision-free, including break-away postures
    assert res["legal"].all()
    # (fingers that miss the mug stop against each other instead: only part of the poses touch the object)
    assert (res["link_dist"] > 0).all() and ((res["link_dist"] < 0.1).sum(axis=1) >= 1).mean() > 0.3
    # joint-level FK equals DOF-level FK wherever the joints still equal ratio x DOF
    r = sc.robots[0]
    jd = [j for c in r.chains for j in c.joints]
    coupled = np.abs(js[:, 7:] - closed["dof"][:, [j.dof for j in jd]] * np.array([j.ratio for j in jd])).max(axis=1) < 1e-12
    raw2 = closed["raw"].copy(); raw2[:, 7:] = closed["dof"].astype(np.float32)
    res2 = eng.eval_batch(rid, obj, raw2[coupled], body_tf=True)
    assert coupled.sum() > 20 and (~coupled).sum() > 5
    assert np.abs(res2["body_tf"] - res["body_tf"][coupled]).max() < 2e-4      # fp32 rounding of the DOF values only


def test_contacts_batch_matches_single_pose_path_and_reference(closed):
    sc, eng, sb, rid, js = closed["sc"], closed["eng"], closed["sb"], closed["rid"], closed["js"]
    osc = OracleScene(sc, nthreads=1)
    ag = AutoGraspOracle(sc, 0, osc.w, osc.ids)
    res = eng.contacts_batch(rid, js, THR, input_mode=B2C_INPUT_JOINTS)
    assert res["offsets"][-1] == len(res["contacts"]) and (np.diff(res["offsets"]) >= 0).all()
    assert (np.diff(res["pose"]) >= 0).all()
    pairs = eng.robot_pairs(rid)
    inv_e = {v: k for k, v in sb.body_ids.items()}
    inv_o = {v: k for k, v in osc.ids.items()}
    got = _by_pose_pair(res, pairs, inv_e)
    bodies = eng.robot_bodies(rid)
    tf = eng.eval_batch(rid, -1, js, input_mode=B2C_INPUT_JOINTS, body_tf=True)["body_tf"]
    interest_o = [osc.ids[inv_e[b]] for b in bodies]
    npairs = exact = nsingle = same_single = 0
    for p in range(0, len(js), 3):
        # (1) the single-pose interface on the same link poses (which it re-derives from q, t: 1e-14 apart) returns the same contacts
        for k, b in enumerate(bodies):
            eng.set_body_transform(b, tf[p, k, :4], tf[p, k, 4:])
        single = eng.all_contacts(THR, interest=bodies)
        want = {}
        for (a, b), cs in single:
            want[tuple(sorted((inv_e[a], inv_e[b])))] = np.array(cs).reshape(-1, 13)
        mine = {k[1]: np.array([c for _, c in v]) for k, v in got.items() if k[0] == p}
        assert set(mine) == set(want)
        for key in want:
            nsingle += 1
            if mine[key].shape == want[key].shape and np.allclose(mine[key], want[key], rtol=1e-9, atol=1e-9):
                same_single += 1
            else:
                # flat-on-flat contacts tie exactly; which of the tied candidates survives the 3 mm duplicate rule then
                # hangs on the last bit of the transforms: the clusters must still agree
                for r in want[key]:
                    assert min(np.linalg.norm(mine[key][:, :3] - r[:3], axis=1)) < 6.0, (p, key)
        # (2) the reference backend at the same joint values: same contacting pairs, same clusters
        ag.set_joints((js[p, 0:4], js[p, 4:7]), js[p, 7:])
        for k, b in enumerate(bodies):      # joint-level FK of the engine == the restated fwdKinematics feeding the reference
            qo, to = ag.last_tf[inv_e[b]]
            assert np.abs(tf[p, k, 4:] - to).max() < 1e-9, (p, k, tf[p, k, 4:] - to)
            assert min(np.abs(tf[p, k, :4] - qo).max(), np.abs(tf[p, k, :4] + qo).max()) < 1e-12, (p, k)
        ref_pairs = {}
        for (a, b), cs in osc.w.all_contacts(THR, interest=interest_o):
            ref_pairs[tuple(sorted((inv_o[a], inv_o[b])))] = (inv_o[a] > inv_o[b], np.array(cs).reshape(-1, 13))
        assert set(ref_pairs) == set(mine), (p, set(ref_pairs), set(mine))
        for key, (swapped_o, co) in ref_pairs.items():
            swapped_e = got[(p, key)][0][0]
            ce = mine[key]
            if swapped_o != swapped_e:
                co = co[:, [3, 4, 5, 0, 1, 2, 9, 10, 11, 6, 7, 8, 12]]
            npairs += 1
            for r in co:
                assert min(np.linalg.norm(ce[:, :3] - r[:3], axis=1)) < 6.0      # same cluster (3 mm duplicate rule)
            if len(ce) == len(co) and {tuple(np.round(r[:6], 5)) for r in ce} == {tuple(np.round(r[:6], 5)) for r in co}:
                exact += 1
    assert npairs >= 30 and exact >= 0.7 * npairs
    assert same_single >= 0.9 * nsingle
    # raw candidates are a superset of the processed contacts
    raw = eng.contacts_batch(rid, js[:8], THR, input_mode=B2C_INPUT_JOINTS, raw=True)
    assert len(raw["contacts"]) >= res["offsets"][8]


def test_chunked_pipeline_equals_one_chunk_calls(closed):
    """b2c_contacts_batch sends a large batch through in chunks of poses, post-processing chunk c on the host cores while the
    GPU works on chunk c + 1 (two pinned staging buffers): the result must be the concatenation of small one-chunk calls."""
    eng, rid, js = closed["eng"], closed["rid"], closed["js"]
    reps = 4500 // len(js) + 1
    big = np.tile(js, (reps, 1))[:4500]                       # 4 500 poses -> chunks of 1 125
    whole = eng.contacts_batch(rid, big, THR, input_mode=B2C_INPUT_JOINTS)
    assert len(whole["contacts"]) > 1000
    parts, at = [], 0
    for lo in range(0, len(big), 700):                        # 700 < 1 024: always a single chunk
        r = eng.contacts_batch(rid, big[lo:lo + 700], THR, input_mode=B2C_INPUT_JOINTS)
        parts.append((r["contacts"], r["pose"] + lo, r["pair"]))
    c = np.concatenate([p[0] for p in parts]); ps = np.concatenate([p[1] for p in parts]); pr = np.concatenate([p[2] for p in parts])
    assert np.array_equal(whole["contacts"], c) and np.array_equal(whole["pose"], ps) and np.array_equal(whole["pair"], pr)
    off = whole["offsets"]
    assert off[0] == 0 and off[-1] == len(c) and (np.diff(off) >= 0).all()
    # every repetition of the fixture's poses reports the same contacts
    n0 = len(js)
    first = whole["contacts"][(whole["pose"] < n0)]
    second = whole["contacts"][(whole["pose"] >= n0) & (whole["pose"] < 2 * n0)]
    assert np.array_equal(first, second)


def test_device_visit_keys_equal_host_replay(closed):
    """The candidates' visit order comes from per-candidate keys computed on the device (csrc/contact_order.cu); the host
    replay of the reference's recursion (B2C_CONTACT_ORDER=host), which tests/test_contact_order.py holds to the reference's
    own contact trace, must give the same contacts, bit for bit -- raw (order itself) and post-processed."""
    eng, rid, js = closed["eng"], closed["rid"], closed["js"]
    big = np.tile(js, (3000 // len(js) + 1, 1))[:3000]
    out = {}
    for mode in ("device", "host"):
        if mode == "host":
            os.environ["B2C_CONTACT_ORDER"] = "host"
        else:
            os.environ.pop("B2C_CONTACT_ORDER", None)
        try:
            out[mode] = (eng.contacts_batch(rid, big, THR, input_mode=B2C_INPUT_JOINTS, raw=True),
                         eng.contacts_batch(rid, big, THR, input_mode=B2C_INPUT_JOINTS))
        finally:
            os.environ.pop("B2C_CONTACT_ORDER", None)
    for i in (0, 1):
        a, b = out["device"][i], out["host"][i]
        assert len(a["contacts"]) == len(b["contacts"]) > 1000
        assert np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["pair"], b["pair"])
        assert np.array_equal(a["contacts"], b["contacts"])


def test_fp32_short_cuts_change_nothing(closed):
    """contact_kernel's two fp32 short cuts -- the root-box cull pass that builds the warps' work list, and the
    separating-axis gate in front of the fp64 feature enumeration -- may only drop work that would have emitted nothing:
    with B2C_CONTACT_EXHAUSTIVE=1 (every query a warp, every queued triangle pair enumerated) the candidates themselves
    (raw, as a set per (pose, pair)) and the processed contacts must be the same, bit for bit."""
    eng, rid, js = closed["eng"], closed["rid"], closed["js"]
    big = np.tile(js, (3000 // len(js) + 1, 1))[:3000]
    out = {}
    for mode in ("fast", "exhaustive"):
        if mode == "exhaustive":
            os.environ["B2C_CONTACT_EXHAUSTIVE"] = "1"
        else:
            os.environ.pop("B2C_CONTACT_EXHAUSTIVE", None)
        try:
            out[mode] = (eng.contacts_batch(rid, big, THR, input_mode=B2C_INPUT_JOINTS, raw=True),
                         eng.contacts_batch(rid, big, THR, input_mode=B2C_INPUT_JOINTS))
        finally:
            os.environ.pop("B2C_CONTACT_EXHAUSTIVE", None)
    for i in (0, 1):
        a, b = out["fast"][i], out["exhaustive"][i]
        assert len(a["contacts"]) == len(b["contacts"]) > 1000
        assert np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["pair"], b["pair"])
        assert np.array_equal(a["contacts"], b["contacts"])


def test_region_batch_matches_single_queries_and_reference():
    """neighbourhoods of every contact of a batch of closed human hands (elastic fingertips): the batched bodyRegion returns
    exactly what one b2c_region call per contact returns, which equals the reference's bodyRegion as a set"""
    from graspit_b200 import engine as E
    from tests.common import aa_to_raw_states, sample_aa_states
    sc = load_pack("humanhand_flask")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    osc = OracleScene(sc, nthreads=1)
    rid, obj = sb.robot_ids[0], sb.body_ids[object_body(sc)]
    r = sc.robots[0]
    pos, _ = sample_aa_states(sc, 400, seed=41, strata=("near",))
    rest = np.clip(r.dof_default, r.dof_min, r.dof_max)
    raw = aa_to_raw_states(sc, pos, np.tile(rest[None, :], (len(pos), 1)).astype(np.float32))
    raw = raw[eng.eval_batch(rid, obj, raw)["legal"] == 1][:24]
    ag = eng.auto_grasp(rid, r, raw)
    ok = (ag["status"] & AG_INTERP_ERROR) == 0
    js = np.concatenate([raw[ok, :7].astype(np.float64), ag["joints"][ok]], axis=1)
    res = eng.contacts_batch(rid, js, THR, input_mode=B2C_INPUT_JOINTS)
    assert len(res["contacts"]) > 40
    pairs = eng.robot_pairs(rid)
    inv = {v: k for k, v in sb.body_ids.items()}
    bodies, points, normals, radii = [], [], [], []
    for c, q in zip(res["contacts"], res["pair"]):
        a, b = pairs[q]
        if max(sc.bodies[inv[a]].youngs, sc.bodies[inv[b]].youngs) <= 0:
            continue                      # rigid pair: addContacts takes the point-contact branch, no neighbourhoods
        rad = E.soft_neighborhood_radius(sc.bodies[inv[a]].youngs, sc.bodies[inv[b]].youngs)
        bodies += [a, b]; points += [c[0:3], c[3:6]]; normals += [c[6:9], c[9:12]]; radii += [rad, rad]
    got = eng.body_region_batch(bodies, points, normals, radii)
    assert len(got) == len(bodies) and len(bodies) >= 20
    nonempty = 0
    for i in range(0, len(bodies), 7):
        one = eng.body_region(bodies[i], points[i], normals[i], radii[i])
        assert np.array_equal(got[i], one), i
        ref_pts = osc.w.region(osc.ids[inv[bodies[i]]], points[i], normals[i], radii[i])
        assert {tuple(x) for x in ref_pts} == {tuple(x) for x in got[i]}
        nonempty += len(one) > 1
    assert nonempty > 5
    # the last point of every list is the query point itself
    assert all(np.array_equal(g[-1], np.asarray(p, dtype=np.float64)) for g, p in zip(got, points))


def test_region_batch_edge_cases(closed):
    """queries that flag nothing (a radius of a micron away from the surface: the list is the query point alone), queries on
    different bodies mixed in one call, and a caller buffer smaller than the result (the lists are cut at `cap` points, the
    offsets and the total still describe the whole result) -- each against the single-query call"""
    import ctypes as C
    eng, sb, sc = closed["eng"], closed["sb"], closed["sc"]
    ids = sorted(sb.body_ids.values())
    rng = np.random.default_rng(5)
    bodies, points, normals, radii = [], [], [], []
    for k in range(60):
        b = ids[k % len(ids)]
        tri = np.asarray(sc.bodies[{v: kk for kk, v in sb.body_ids.items()}[b]].tris, dtype=np.float64).reshape(-1, 3, 3)
        t = tri[rng.integers(len(tri))]
        n = np.cross(t[1] - t[0], t[2] - t[0]); n /= max(np.linalg.norm(n), 1e-30)
        far = k % 3 == 0
        p = t.mean(axis=0) + (500.0 * n if far else 0.0)
        bodies.append(b); points.append(p); normals.append(-n); radii.append(1.0e-3 if far else 6.0)
    got = eng.body_region_batch(bodies, points, normals, radii)
    assert len(got) == 60
    for i in range(60):
        one = eng.body_region(bodies[i], points[i], normals[i], radii[i])
        assert np.array_equal(got[i], one), i
        if i % 3 == 0:
            assert len(got[i]) == 1 and np.array_equal(got[i][0], points[i])
    assert sum(len(g) > 3 for g in got) > 20
    # a buffer that holds only part of the result
    total = sum(len(g) for g in got)
    cap = total // 2
    b = np.ascontiguousarray(bodies, dtype=np.int32); p = np.ascontiguousarray(points); nn = np.ascontiguousarray(normals)
    r = np.ascontiguousarray(radii, dtype=np.float64)
    out = np.full((cap, 3), np.nan); off = np.zeros(61, dtype=np.int32); tot = C.c_int(0)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    eng._ck(eng.L.b2c_region_batch(eng.h, 60, b.ctypes.data_as(C.POINTER(C.c_int)), dp(p), dp(nn), dp(r), dp(out), cap,
                                   off.ctypes.data_as(C.POINTER(C.c_int)), C.byref(tot)))
    assert tot.value == total and off[60] == total
    flat = np.concatenate(got)
    assert np.array_equal(off[:60], np.cumsum([0] + [len(g) for g in got])[:60])
    assert np.array_equal(out, flat[:cap])


def test_contacts_batch_without_any_contact(closed):
    """hands far away from everything: no candidates, no groups, all offsets zero (the empty batch goes through the same
    chunked pipeline)"""
    eng, rid, js = closed["eng"], closed["rid"], closed["js"]
    far = np.tile(js[:1], (2500, 1))
    far[:, 4:7] += np.array([5000.0, 0.0, 0.0])             # translation of the base: 5 m away from the object
    far[:, 7:] = 0.0                                          # open hand (no self contact)
    res = eng.contacts_batch(rid, far, THR, input_mode=B2C_INPUT_JOINTS)
    assert len(res["contacts"]) == 0 and len(res["pose"]) == 0
    assert res["offsets"].shape == (2501,) and not res["offsets"].any()
    # ... and a batch where only a few poses touch anything
    mixed = far.copy(); mixed[::500] = js[:5]
    res = eng.contacts_batch(rid, mixed, THR, input_mode=B2C_INPUT_JOINTS)
    one = eng.contacts_batch(rid, js[:5], THR, input_mode=B2C_INPUT_JOINTS)
    assert len(res["contacts"]) == len(one["contacts"]) > 0
    assert np.array_equal(res["contacts"], one["contacts"]) and np.array_equal(res["pose"] // 500, one["pose"])
