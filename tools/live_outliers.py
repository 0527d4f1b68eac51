"""Diagnostic (GPU box): LIVE poses whose energy or link distances differ from the reference by more than 1e-4 relative,
with the per-link distances and closest points of both sides (single-pose interface on the reference's link poses)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graspit_b200.engine import B2C_INPUT_AA_EIGEN, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import fk as ofk
from oracle.port import transf_np as T
from tests.common import OracleScene, object_body, sample_aa_states
npre, nlive, seed = 100000, 60000, 20261020
sc = load_pack("barrett_mug"); obj = object_body(sc); r = sc.robots[0]
eng = Engine(0); sb = SceneBinding(eng, sc)
osc = OracleScene(sc, nthreads=ref.hardware_threads())
rng = np.random.default_rng(seed)
pos, _ = sample_aa_states(sc, npre, seed=seed, strata=("far", "near", "close"))
states = np.concatenate([pos, rng.uniform(-4, 4, size=(npre, r.eg.shape[0]))], 1).astype(np.float32)[:nlive]
s64 = states.astype(np.float64)
base = ofk.aa_state_to_base(r, sc.body_tran[obj], s64[:, :6])
raw64 = np.concatenate([base[0], base[1], ofk.eigen_to_dof(r, s64[:, 6:])], 1)
ll, el, dl, order, tf7 = osc.eval(0, obj, raw64, mode=1, want_dist=True)
live = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], states, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[obj], live=True, link_dist=True)
m = (ll == 1) & (live["legal"] == 1)
rel = np.abs(live["energy"] - el) / np.maximum(np.abs(el), 1e-9)
ld = live["link_dist"].astype(np.float64); both = (dl >= 0) & (ld >= 0)
drel = np.where(both, np.abs(ld - dl) / np.maximum(dl, 1e-9), 0)
print("link-dist outliers (pose, link):", np.argwhere(drel > 1e-4).tolist(), drel[drel > 1e-4])
bad = np.unique(np.argwhere(drel > 1e-4)[:, 0])
print("outliers", bad, rel[bad])
w = osc.w
for p in bad:
    print("pose", p, "E ours", live["energy"][p], "ref", el[p])
    for j, b in enumerate(order):
        w.set_transform(osc.ids[b], tf7[p, j, :4], tf7[p, j, 4:]); eng.set_body_transform(sb.body_ids[b], tf7[p, j, :4], tf7[p, j, 4:])
    for j, b in enumerate(order):
        dr, p1r, p2r = w.body_distance(osc.ids[b], osc.ids[obj])
        de, p1e, p2e = eng.body_to_body_distance(sb.body_ids[b], sb.body_ids[obj])[:3]
        tl = T.make(tf7[p, j, :4][None], tf7[p, j, 4:][None]); to = T.make(sc.body_tran[obj][0][None], sc.body_tran[obj][1][None])
        d1 = np.linalg.norm(np.asarray(p1r) - np.asarray(p1e)); d2 = np.linalg.norm(np.asarray(p2r) - np.asarray(p2e))
        flag = " <--" if (max(d1, d2) > 1e-3 or drel[p, j] > 1e-4) else ""
        print("  link %d d ref %.9f ours %.9f rel %.2e | batch link_dist %.9f | |dp1| %.4g |dp2| %.4g%s" % (j, dr, de, abs(de - dr) / max(dr, 1e-9), live["link_dist"][p, j], d1, d2, flag))
