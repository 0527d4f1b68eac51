"""K6 / K7 parity: contact candidates, contact(), allContacts() and bodyRegion() against the reference backend."""
import numpy as np
import pytest

from graspit_b200.engine import Engine
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import transf_np as T

pytestmark = pytest.mark.gpu
THR = 0.2      # Contact::THRESHOLD (src/contact/contact.cpp:58)


def _rows(a, nd=6):
    return {tuple(np.round(r[:6], nd)) for r in a}


def _near_poses(w, mug, palm, n, seed=7):
    """palm poses 0.1 mm away from the mug, built from the reference's own closest points"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.uniform(-150, 150, 3) + np.array([0, 0, 105.0])
        w.set_transform(mug, [1, 0, 0, 0], [0, 0, 0]); w.set_transform(palm, q, t)
        d, p1, p2 = w.body_distance(palm, mug)
        if d <= 0:
            continue
        tr = T.make(q[None], t[None])
        pw1 = T.apply(tr, T.apply(tr, p1[None]))[0]
        dirv = (p2 - pw1) / np.linalg.norm(p2 - pw1)
        out.append((q, t + dirv * (d - 0.1)))
    return out


@pytest.fixture(scope="module")
def worlds():
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    w = ref.RefWorld()
    ids_e = [eng.add_body(sc.bodies[i].tris) for i in (0, 1)]
    ids_o = [w.add_body(sc.bodies[i].tris.astype(np.float64)) for i in (0, 1)]
    return sc, eng, w, ids_e, ids_o


def test_contact_candidates_and_processed_sets(worlds):
    sc, eng, w, (mug_e, palm_e), (mug_o, palm_o) = worlds
    exact = same_order = 0
    poses = _near_poses(w, mug_o, palm_o, 30)
    for q, t in poses:
        w.set_transform(palm_o, q, t); w.set_transform(mug_o, [1, 0, 0, 0], [0, 0, 0])
        eng.set_body_transform(palm_e, q, t); eng.set_body_transform(mug_e, [1, 0, 0, 0], [0, 0, 0])
        raw_o, _ = w.contact_raw(palm_o, mug_o, THR)
        raw_e = eng.contact(palm_e, mug_e, THR, raw=True)
        assert len(raw_o) > 0
        # candidate point pairs: equal as sets (the reference may hold endpoint duplicates, Appendix A.16)
        assert _rows(raw_e) == _rows(raw_o)
        # ... and in the reference's own visit order (csrc/contacts_host.cpp obb_visit_order): compare the sequences of
        # first occurrences (the reference's endpoint duplicates repeat a row, they never introduce one)
        def firsts(a):
            seen, out = set(), []
            for r in a:
                k = tuple(np.round(r[:6], 6))
                if k not in seen:
                    seen.add(k); out.append(k)
            return out
        same_order += firsts(raw_e) == firsts(raw_o)
        # every candidate obeys the reference's conventions
        for r in raw_e:
            assert np.linalg.norm(r[6:9]) == pytest.approx(1.0, abs=1e-9)
        full_o = w.contact(palm_o, mug_o, THR)
        full_e = eng.contact(palm_e, mug_e, THR)
        assert len(full_e) >= 1
        # processed contact sets: the 3 mm duplicate removal and the perimeter compaction are order dependent, the
        # candidates arrive in the reference's order, so the sets are the reference's
        for r in full_o:
            assert min(np.linalg.norm(full_e[:, :3] - r[:3], axis=1)) <= 3.0
        for r in full_e:
            assert min(np.linalg.norm(full_o[:, :3] - r[:3], axis=1)) <= 3.0
        if len(full_e) == len(full_o) and _rows(full_e, 5) == _rows(full_o, 5):
            exact += 1
    print("contact(): identical processed sets", exact, "of", len(poses), "; candidate sequences identical", same_order)
    assert same_order >= 0.9 * len(poses)
    assert exact >= 0.97 * len(poses)


def test_no_contacts_when_far_or_interpenetrating(worlds):
    sc, eng, w, (mug_e, palm_e), (mug_o, palm_o) = worlds
    eng.set_body_transform(mug_e, [1, 0, 0, 0], [0, 0, 0])
    eng.set_body_transform(palm_e, [1, 0, 0, 0], [0, 0, 400.0])
    assert len(eng.contact(palm_e, mug_e, THR)) == 0
    assert eng.all_contacts(THR) == []


def test_all_contacts_pair_sets(worlds):
    """allContacts over a whole hand: the set of contacting body pairs is exact."""
    sc = load_pack("barrett_mug")
    from graspit_b200.engine import SceneBinding
    from tests.common import OracleScene, aa_to_raw_states, object_body, sample_aa_states
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    osc = OracleScene(sc)
    pos, dofs = sample_aa_states(sc, 300, seed=21, strata=("close",))
    raw = aa_to_raw_states(sc, pos, dofs)
    order, tf7 = osc.link_poses(0, raw)
    inv_o = {v: k for k, v in osc.ids.items()}
    inv_e = {v: k for k, v in sb.body_ids.items()}
    npairs = 0
    for p in range(0, 300, 3):
        for k, b in enumerate(order):
            osc.w.set_transform(osc.ids[b], tf7[p, k, :4], tf7[p, k, 4:])
            eng.set_body_transform(sb.body_ids[b], tf7[p, k, :4], tf7[p, k, 4:])
        co = osc.w.all_contacts(THR)
        ce = eng.all_contacts(THR)
        so = {tuple(sorted((inv_o[a], inv_o[b]))) for (a, b), _ in co}
        se = {tuple(sorted((inv_e[a], inv_e[b]))) for (a, b), _ in ce}
        assert so == se
        npairs += len(so)
    assert npairs > 5


def test_body_region(worlds):
    """bodyRegion (A10): vertices within the radius, normal filter, reference point appended."""
    sc, eng, w, (mug_e, palm_e), (mug_o, palm_o) = worlds
    rng = np.random.default_rng(3)
    tris = sc.bodies[0].tris.astype(np.float64)
    total = 0
    for trial in range(25):
        t = tris[rng.integers(len(tris))]
        p = t.mean(axis=0) + rng.normal(size=3) * 0.05
        n = np.cross(t[1] - t[0], t[2] - t[0]); n /= np.linalg.norm(n)
        n = -n                                   # contact normals point inwards
        radius = [0.5, 3.5, 8.0][trial % 3]
        ro = w.region(mug_o, p, n, radius)
        re = eng.body_region(mug_e, p, n, radius)
        assert np.allclose(re[-1], p) and np.allclose(ro[-1], p)
        assert {tuple(r) for r in ro} == {tuple(r) for r in re}
        assert len(ro) == len(re)
        total += len(re)
    assert total > 100
