"""BASELINE.json configs 3-5 as parity cases: HumanHand20DOF vs flask (contacts), Staubli arm +
Barrett hand vs obstacles (full-robot collision), Barrett vs the 220 800-triangle mug.  Oracle = the reference's
own GraspitCollision; sizes chosen so the CPU side finishes in seconds."""
import numpy as np
import pytest

from graspit_b200.engine import Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import transf_np as T
from tests.common import OracleScene, aa_to_raw_states, hand_robot, object_body, sample_aa_states

pytestmark = pytest.mark.gpu
REL = 1e-4
THR = 0.2


def relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-9)


def _pair_sets_engine(eng, sb, rid, res, n):
    pairs = eng.robot_pairs(rid)
    inv = {v: k for k, v in sb.body_ids.items()}
    out = []
    for p in range(n):
        out.append({tuple(sorted((inv[a], inv[b]))) for i, (a, b) in enumerate(pairs)
                    if (int(res["pair_mask"][p, i >> 6]) >> (i & 63)) & 1})
    return out


# ------------------------------------------------------------------------------------------------ config 3
@pytest.fixture(scope="module")
def human():
    sc = load_pack("humanhand_flask")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    osc = OracleScene(sc, nthreads=min(8, ref.hardware_threads()))
    return dict(sc=sc, eng=eng, sb=sb, osc=osc, obj=object_body(sc), rid=sb.robot_ids[0])


def _human_states(sc, n, seed, spread):
    """planner-box / 150 mm-ball positions; DOFs drawn around the rest posture (fully random 20-DOF postures
    self-collide almost always)"""
    r = sc.robots[0]
    pos, dofs = sample_aa_states(sc, n, seed=seed, strata=("far", "near"))
    mid = np.clip(r.dof_default, r.dof_min, r.dof_max)
    d = mid[None, :] + spread * (dofs.astype(np.float64) - mid[None, :])
    return aa_to_raw_states(sc, pos, d.astype(np.float32))


def test_config3_humanhand_legality_energy_distances(human):
    """20-DOF hand, 16 bodies, ~16k triangles: FK, legality, PRESET energy and per-link distances."""
    sc, eng, sb, osc = human["sc"], human["eng"], human["sb"], human["osc"]
    raw = _human_states(sc, 300, 77, 0.5)
    legal_o, energy_o, dist_o, order, tf7 = osc.eval(0, human["obj"], raw, mode=0, want_dist=True)
    res = eng.eval_batch(human["rid"], sb.body_ids[human["obj"]], raw, pair_mask=True, link_dist=True, body_tf=True)
    assert np.abs(res["body_tf"][:, :, 4:7] - tf7[:, :, 4:7]).max() < 1e-8
    assert (res["legal"] == legal_o).all()
    assert 20 < legal_o.sum() < 290
    m = legal_o == 1
    assert relerr(res["energy"][m], energy_o[m]).max() < REL
    ld = res["link_dist"].astype(np.float64)
    assert ((ld < 0) == (dist_o < 0)).all()
    both = dist_o >= 0
    assert relerr(ld[both], dist_o[both]).max() < REL
    want = osc.pair_sets(0, raw[:100])
    got = _pair_sets_engine(eng, sb, human["rid"], res, 100)
    assert got == want


def _canon(a, b, contacts, inv):
    """contacts of a pair with body 1 = the lower scene index (the reference orders a pair by model address)"""
    c = np.array(contacts, dtype=np.float64).reshape(-1, 13)
    ia, ib = inv[a], inv[b]
    if ia > ib:
        c = c[:, [3, 4, 5, 0, 1, 2, 9, 10, 11, 6, 7, 8, 12]]
        ia, ib = ib, ia
    return (ia, ib), c


def _fit_cond(pos, normal, pts):
    from oracle.port import contact_np as cn
    if len(pts) < 3:
        return 1.0
    _, _, aff, _ = cn.contact_frame(pos, normal)
    p = np.asarray(pts) @ aff
    B = np.stack([p[:, 0] ** 2, p[:, 1] ** 2, p[:, 0] * p[:, 1]], axis=1)
    return float(np.linalg.cond(B.T @ B))


def test_config3_allcontacts_with_region(human):
    """allContacts(0.2) over the whole hand for poses pushed into contact, + bodyRegion around each contact
    (the soft-finger neighbourhood radius of contactSetting.cpp:129 for E = 1.5e6: 3.5 mm)."""
    sc, eng, sb, osc = human["sc"], human["eng"], human["sb"], human["osc"]
    obj = human["obj"]
    raw = _human_states(sc, 300, 78, 0.3)
    legal_o, _, dist_o, order, tf7 = osc.eval(0, obj, raw, mode=0, want_dist=True)
    inv_o = {v: k for k, v in osc.ids.items()}
    inv_e = {v: k for k, v in sb.body_ids.items()}
    w = osc.w
    done = npairs = ncontacts = exact = ncones = nsoft = 0
    for p in np.nonzero(legal_o)[0]:
        if done >= 25:
            break
        d = dist_o[p]
        k = int(np.argmin(d))
        if d[k] <= THR:
            continue
        # move the whole hand towards the object along the closest-point direction of its nearest link
        for j, b in enumerate(order):
            w.set_transform(osc.ids[b], tf7[p, j, :4], tf7[p, j, 4:])
        w.set_transform(osc.ids[obj], *sc.body_tran[obj])
        dd, p1, p2 = w.body_distance(osc.ids[order[k]], osc.ids[obj])
        tl = T.make(tf7[p, k, :4][None], tf7[p, k, 4:][None])
        to = T.make(sc.body_tran[obj][0][None], sc.body_tran[obj][1][None])
        w1 = T.apply(tl, T.apply(tl, p1[None]))[0]       # undo the reference's double inverse
        w2 = T.apply(to, T.apply(to, p2[None]))[0]
        shift = (w2 - w1) / np.linalg.norm(w2 - w1) * (dd - 0.1)
        for j, b in enumerate(order):
            w.set_transform(osc.ids[b], tf7[p, j, :4], tf7[p, j, 4:] + shift)
            eng.set_body_transform(sb.body_ids[b], tf7[p, j, :4], tf7[p, j, 4:] + shift)
        ncol, _ = w.all_collisions(fast=True)
        if ncol:
            continue
        co = w.all_contacts(THR)
        ce = eng.all_contacts(THR)
        so = {tuple(sorted((inv_o[a], inv_o[b]))) for (a, b), _ in co}
        se = {tuple(sorted((inv_e[a], inv_e[b]))) for (a, b), _ in ce}
        assert so == se and len(so) >= 1
        done += 1
        npairs += len(so)
        by_pair_o = dict(_canon(a, b, c, inv_o) for (a, b), c in co)
        by_pair_e = dict(_canon(a, b, c, inv_e) for (a, b), c in ce)
        for key, cl in by_pair_o.items():
            ec = by_pair_e[key]
            ncontacts += len(cl)
            # clusters agree: the 3 mm duplicate removal keeps a representative that depends on candidate order, so
            # two valid representatives of one cluster are < 2 x 3 mm apart; most sets agree exactly
            for r in cl:
                assert min(np.linalg.norm(ec[:, :3] - r[:3], axis=1)) < 6.0
            if len(ec) == len(cl) and {tuple(np.round(r[:6], 5)) for r in ec} == {tuple(np.round(r[:6], 5)) for r in cl}:
                exact += 1
                # friction-cone setup per contact (addContacts: checkContactNormals -> PointContact frames + cone wrenches)
                from graspit_b200 import engine as E
                from oracle.port import contact_np as cn
                ta = (tf7[p, order.index(key[0]), :4], tf7[p, order.index(key[0]), 4:] + shift) if key[0] in order else sc.body_tran[key[0]]
                tb = (tf7[p, order.index(key[1]), :4], tf7[p, order.index(key[1]), 4:] + shift) if key[1] in order else sc.body_tran[key[1]]
                fixed, _ = E.check_contact_normals(ec, ta, tb)
                wr = E.point_contact_wrenches(fixed, 1, 0.5, [0.0, 0.0, 0.0], 100.0)
                fr = E.contact_frames(fixed, 1)
                for i, row in enumerate(fixed):
                    rr, _ok = cn.check_contact_normals(ec[i], ta, tb)
                    assert np.allclose(rr, row)
                    assert np.allclose(wr[i], cn.point_contact_wrenches(row[0:3], row[6:9], 0.5, [0.0, 0.0, 0.0], 100.0), atol=1e-10)
                    assert np.allclose(fr[i, 4:], row[0:3])
                ncones += len(fixed)
                # elastic fingertips: the soft branch of addContacts -- neighbourhoods of both bodies around every
                # contact (device region kernel), merge, paraboloid fit, contact ellipse, 20-edge ellipsoid wrenches
                ya, yb = sc.bodies[key[0]].youngs, sc.bodies[key[1]].youngs
                if max(ya, yb) > 0:
                    rad = E.soft_neighborhood_radius(ya, yb)
                    assert abs(rad - cn.soft_neighborhood_radius(ya, yb)) < 1e-12
                    n1e = [eng.body_region(sb.body_ids[key[0]], r[0:3], r[6:9], rad) for r in fixed]
                    n2e = [eng.body_region(sb.body_ids[key[1]], r[3:6], r[9:12], rad) for r in fixed]
                    for r, a1, a2 in zip(fixed, n1e, n2e):
                        assert {tuple(x) for x in w.region(osc.ids[key[0]], r[0:3], r[6:9], rad)} == {tuple(x) for x in a1}
                        assert {tuple(x) for x in w.region(osc.ids[key[1]], r[3:6], r[9:12], rad)} == {tuple(x) for x in a2}
                    mr, m1, m2, nm = E.soft_contact_merge(fixed, ta, tb, n1e, n2e)
                    mr_o, m1_o, m2_o, nm_o = cn.merge_soft_neighborhoods(fixed, ta, tb, n1e, n2e)
                    assert nm == nm_o and np.array_equal(mr, mr_o)
                    assert all(np.array_equal(x, y) for x, y in zip(m1 + m2, m1_o + m2_o))
                    f1 = E.soft_contact_fit(mr, 1, m1); f2 = E.soft_contact_fit(mr, 2, m2)
                    wr1, el1 = E.soft_contact_wrenches(mr, 1, f1, f2, ta[0], tb[0], max(ya, yb), 1.0, [0.0, 0.0, 0.0], 100.0)
                    for i, row in enumerate(mr):
                        fo1, _ = ref.soft_fit(row[0:3], row[6:9], m1[i]); fo2, _ = ref.soft_fit(row[3:6], row[9:12], m2[i])
                        # un-recentred neighbourhoods => ill-conditioned normal equations (tests/test_soft_contacts.py)
                        tol = max(1e-9, 1e-13 * max(_fit_cond(row[0:3], row[6:9], m1[i]), _fit_cond(row[3:6], row[9:12], m2[i])))
                        if tol > 1e-3:
                            continue
                        assert np.allclose(f1[i], fo1, rtol=tol, atol=1e-12, equal_nan=True)
                        assert np.allclose(f2[i], fo2, rtol=tol, atol=1e-12, equal_nan=True)
                        sa = cn.soft_contact_side(row[0:3], row[6:9], m1[i], ta[0]); sb_ = cn.soft_contact_side(row[3:6], row[9:12], m2[i], tb[0])
                        wo, eo = cn.soft_contact_wrenches(sa, sb_, max(ya, yb), 1.0, [0.0, 0.0, 0.0], 100.0)
                        if np.isfinite(eo).all() and np.isfinite(el1[i]).all():
                            assert np.allclose(el1[i], eo, rtol=max(1e-6, 100 * tol), atol=1e-9)
                            assert np.allclose(wr1[i], wo, rtol=max(1e-6, 100 * tol), atol=1e-7)
                            nsoft += 1
            # region around the first contact on the object side (elastic fingertip neighbourhoods)
            a_is_obj = key[0] == obj
            cpos = cl[0][0:3] if a_is_obj else cl[0][3:6]
            cnrm = cl[0][6:9] if a_is_obj else cl[0][9:12]
            ro = w.region(osc.ids[obj], cpos, cnrm, 3.5)
            re = eng.body_region(sb.body_ids[obj], cpos, cnrm, 3.5)
            assert {tuple(r) for r in ro} == {tuple(r) for r in re}
    assert done >= 10 and npairs >= done and ncontacts >= done
    print('config 3: identical contact sets in', exact, 'of', npairs, 'contacting pairs')
    assert exact >= 0.99 * npairs
    assert ncones >= 5
    print('soft contacts checked:', nsoft)


# ------------------------------------------------------------------------------------------------ config 4
def test_config4_staubli_barrett_full_robot_collisions():
    """arm (6 DOF, ~47k triangles) + attached Barrett hand vs floor / table / shelf: allCollisions(ALL, interest =
    every robot body) for configurations uniform in the joint limits; coordinates reach ~1.1 m."""
    sc = load_pack("staubli_barrett")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    osc = OracleScene(sc, nthreads=1)
    arm, hand = sc.robots[0], sc.robots[1]
    rng = np.random.default_rng(1234)
    n = 240
    base = np.tile(np.concatenate([arm.tran[0], arm.tran[1]])[None, :], (n, 1))
    dofs = np.concatenate([rng.uniform(arm.dof_min, arm.dof_max, size=(n, arm.n_dof)),
                           rng.uniform(hand.dof_min, hand.dof_max, size=(n, hand.n_dof))], 1)
    raw = np.concatenate([base, dofs], 1).astype(np.float32)
    rid = sb.robot_ids[0]
    res = eng.eval_batch(rid, -1, raw, pair_mask=True, body_tf=True)
    order, tf7 = osc.link_poses(0, raw)
    assert [sb.body_ids[b] for b in order] == eng.robot_bodies(rid)
    assert np.abs(res["body_tf"][:, :, 4:7] - tf7[:, :, 4:7]).max() < 1e-7
    want = osc.pair_sets(0, raw)
    got = _pair_sets_engine(eng, sb, rid, res, n)
    assert got == want
    ncol = sum(len(g) for g in got)
    assert ncol > 50
    assert [(len(g) == 0) for g in got] == [bool(x) for x in res["legal"]]


# ------------------------------------------------------------------------------------------------ config 5
def test_config5_200k_triangle_object():
    """Barrett vs the mug subdivided to 220 800 triangles.  Midpoint subdivision does not move the surface, so the
    results must equal those on the 3 450-triangle mug (size-independent property), and a sample is checked
    against the oracle built on the big mesh."""
    small, big = load_pack("barrett_mug"), load_pack("barrett_mug_200k")
    assert len(big.bodies[0].tris) == 220800
    e1, e2 = Engine(0), Engine(0)
    s1, s2 = SceneBinding(e1, small), SceneBinding(e2, big)
    obj = object_body(small)
    pos, dofs = sample_aa_states(small, 20000, seed=4321)
    raw = aa_to_raw_states(small, pos, dofs)
    r1 = e1.eval_batch(s1.robot_ids[0], s1.body_ids[obj], raw, link_dist=True)
    r2 = e2.eval_batch(s2.robot_ids[0], s2.body_ids[obj], raw, link_dist=True)
    # fp32 midpoints move vertices by <= 1 ulp: only poses within ~1e-5 mm of tangency may flip
    flips = np.nonzero(r1["legal"] != r2["legal"])[0]
    assert len(flips) <= 2
    m = (r1["legal"] == 1) & (r2["legal"] == 1)
    assert m.sum() > 5000
    assert relerr(r2["energy"][m], r1["energy"][m]).max() < REL
    d1, d2 = r1["link_dist"][m].astype(np.float64), r2["link_dist"][m].astype(np.float64)
    # fp32 midpoints move the subdivided surface by up to an ulp of the coordinates (1e-5 mm)
    assert (np.abs(d2 - d1) <= REL * d1 + 2e-5).all()
    # LIVE on the big mesh against LIVE on the small one
    l1 = e1.eval_batch(s1.robot_ids[0], s1.body_ids[obj], raw[:4000], live=True)
    l2 = e2.eval_batch(s2.robot_ids[0], s2.body_ids[obj], raw[:4000], live=True)
    mm = (l1["legal"] == 1) & (l2["legal"] == 1)
    assert relerr(l2["energy"][mm], l1["energy"][mm]).max() < 5 * REL     # two fp32 results, each within REL
    # oracle on the 220 800-triangle mesh (its tree build dominates: keep the sample small)
    osc = OracleScene(big, nthreads=min(8, ref.hardware_threads()))
    legal_o, energy_o, _, _, _ = osc.eval(0, obj, raw[:400], mode=0)
    assert (r2["legal"][:400] == legal_o).all()
    k = legal_o == 1
    assert relerr(r2["energy"][:400][k], energy_o[k]).max() < REL
