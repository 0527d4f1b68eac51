"""CONTACT_LIVE over a 60 000-state sweep (the size at which round 1 met 5 energies 1.0-1.8e-4 off the reference's):
every legal pose must agree with the reference to north_star's 1e-4, EXCEPT poses where the two tree searches picked a
different closest triangle pair for some link -- and there an exhaustive minimum over every triangle pair, computed
with the reference's own leaf arithmetic (oracle ref_brute_distance: DistanceCallback::leafTest without the hierarchy,
reference collisionAlgorithms_inl.h:176-204), must side with the engine: engine distance == exhaustive distance <=
reference distance, and the energy gap must vanish once the reference's pair for that link is replaced by the exhaustive
one (ContactEnergy::energy + World::findVirtualContact restated below, contactEnergy.cpp:9-55, world.cpp:1641-1651)."""
import numpy as np
import pytest

from graspit_b200.engine import B2C_INPUT_AA_EIGEN, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import fk as ofk
from oracle.port import transf_np as T
from tests.common import OracleScene, object_body, sample_aa_states

pytestmark = pytest.mark.gpu
TOL = 1.0e-4            # north_star: distances and energies within 1e-4 relative


def _live_energy(w, obj_id, points):
    """mean over the hand's bodies of |v| + 50 (1 - cn . v / |v|): contact location = mP1 read as a world point, contact
    normal = -(mP1 - mP2) / |mP1 - mP2| (the reference's frame quirks), v = closest object point - location"""
    terms = []
    for mp1, mp2 in points:
        n = mp1 - mp2
        n = n / np.linalg.norm(n)
        d, cp, _ = w.point_distance(obj_id, mp1)
        v = cp - mp1
        terms.append(np.linalg.norm(v) + 50.0 * (1.0 - np.dot(-n, v) / np.linalg.norm(v)))
    return float(np.mean(terms))


def test_live_60k_sweep_outliers_are_misses_of_the_reference_search():
    nlive, seed = 60000, 31337
    sc = load_pack("barrett_mug")
    obj = object_body(sc)
    r = sc.robots[0]
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    osc = OracleScene(sc, nthreads=max(1, ref.hardware_threads()))
    rng = np.random.default_rng(seed)
    pos, _ = sample_aa_states(sc, 700000, seed=seed, strata=("far", "near", "close"))       # the round-1 sweep's stream
    states = np.concatenate([pos, rng.uniform(-4, 4, size=(700000, r.eg.shape[0]))], 1).astype(np.float32)[:nlive]
    s64 = states.astype(np.float64)
    base = ofk.aa_state_to_base(r, sc.body_tran[obj], s64[:, :6])
    raw64 = np.concatenate([base[0], base[1], ofk.eigen_to_dof(r, s64[:, 6:])], 1)
    ll, el, dl, order, tf7 = osc.eval(0, obj, raw64, mode=1, want_dist=True)
    live = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], states, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[obj],
                          live=True, link_dist=True)
    assert (ll == live["legal"]).all()
    m = ll == 1
    assert m.sum() > 30000
    ld = live["link_dist"].astype(np.float64)
    assert ((dl < 0) == (ld < 0)).all()
    both = (dl >= 0.05) & (ld >= 0)
    assert (np.abs(ld[both] - dl[both]) / dl[both]).max() < TOL
    small = (dl >= 0) & (dl < 0.05)
    if small.any():
        assert np.abs(ld[small] - dl[small]).max() < 1e-5          # micro-distances: absolute (tangency band of north_star)
    rel = np.abs(live["energy"] - el) / np.maximum(np.abs(el), 1e-9)
    bad = np.nonzero(m & (rel > TOL))[0]
    print("LIVE sweep: %d legal poses, max energy rel err %.3g, %d above %.0e" % (int(m.sum()), float(rel[m].max()), len(bad), TOL))
    assert len(bad) <= 12, "more outliers than near-tie geometry explains"
    w = osc.w
    o_id, o_e = osc.ids[obj], sb.body_ids[obj]
    to = T.make(sc.body_tran[obj][0][None], sc.body_tran[obj][1][None])
    for p in bad:
        for j, b in enumerate(order):
            w.set_transform(osc.ids[b], tf7[p, j, :4], tf7[p, j, 4:])
            eng.set_body_transform(sb.body_ids[b], tf7[p, j, :4], tf7[p, j, 4:])
        w.set_transform(o_id, *sc.body_tran[obj]); eng.set_body_transform(o_e, *sc.body_tran[obj])
        ref_pts, fixed_pts, differing = [], [], 0
        for j, b in enumerate(order):
            tl = T.make(tf7[p, j, :4][None], tf7[p, j, 4:][None])
            dr, p1r, p2r = w.body_distance(osc.ids[b], o_id)
            mp1r, mp2r = T.apply(tl, p1r[None])[0], T.apply(to, p2r[None])[0]        # undo the extra inverse: mP1, mP2
            de, l1, l2 = eng.body_to_body_distance(sb.body_ids[b], o_e, local=True)
            ref_pts.append((mp1r, mp2r))
            if max(np.linalg.norm(l1 - mp1r), np.linalg.norm(l2 - mp2r)) > 1e-3:
                # the two searches chose different pairs: the exhaustive minimum decides
                differing += 1
                db, b1, b2, _ = w.brute_distance(osc.ids[b], o_id)
                assert db <= dr * (1 + 1e-12), (p, j, db, dr)                           # the reference did not find the minimum ...
                assert abs(de - db) <= 1e-9 * max(1.0, db), (p, j, de, db)              # ... the engine did
                assert max(np.linalg.norm(l1 - b1), np.linalg.norm(l2 - b2)) < 1e-6, (p, j)
                assert (dr - db) / db < TOL                                             # and the DISTANCES still agree to tolerance
                fixed_pts.append((b1, b2))
            else:
                assert abs(de - dr) <= 1e-9 * max(1.0, dr), (p, j, de, dr)
                fixed_pts.append((mp1r, mp2r))
        assert differing >= 1, (p, "energy differs although every link found the reference's pair")
        e_ref = _live_energy(w, o_id, ref_pts)
        assert abs(e_ref - el[p]) <= 1e-9 * abs(el[p]), (p, e_ref, el[p])               # the restated caller loop is the oracle's
        e_fix = _live_energy(w, o_id, fixed_pts)
        assert abs(e_fix - live["energy"][p]) <= TOL * abs(e_fix), (p, e_fix, float(live["energy"][p]), float(el[p]))
