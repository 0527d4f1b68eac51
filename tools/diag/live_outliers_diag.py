"""Diagnostic (GPU box): the LIVE outliers of the 60k sweep -- per link the reference's distance, the engine's (single-pose
interface) and the exhaustive minimum over all triangle pairs (reference leaf arithmetic)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graspit_b200.engine import B2C_INPUT_AA_EIGEN, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import fk as ofk
from oracle.port import transf_np as T
from tests.common import OracleScene, object_body, sample_aa_states

nlive, seed = 60000, 31337
sc = load_pack("barrett_mug"); obj = object_body(sc); r = sc.robots[0]
eng = Engine(0); sb = SceneBinding(eng, sc)
osc = OracleScene(sc, nthreads=max(1, ref.hardware_threads()))
rng = np.random.default_rng(seed)
pos, _ = sample_aa_states(sc, 700000, seed=seed, strata=("far", "near", "close"))
states = np.concatenate([pos, rng.uniform(-4, 4, size=(700000, r.eg.shape[0]))], 1).astype(np.float32)[:nlive]
s64 = states.astype(np.float64)
base = ofk.aa_state_to_base(r, sc.body_tran[obj], s64[:, :6])
raw64 = np.concatenate([base[0], base[1], ofk.eigen_to_dof(r, s64[:, 6:])], 1)
ll, el, dl, order, tf7 = osc.eval(0, obj, raw64, mode=1, want_dist=True)
live = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], states, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[obj], live=True, link_dist=True)
m = ll == 1
rel = np.abs(live["energy"] - el) / np.maximum(np.abs(el), 1e-9)
bad = np.nonzero(m & (rel > 1e-4))[0]
print("lib", os.environ.get("B2C_LIB", "default"), "outliers", bad.tolist(), rel[bad])
w = osc.w
to = T.make(sc.body_tran[obj][0][None], sc.body_tran[obj][1][None])
for p in bad:
    for j, b in enumerate(order):
        w.set_transform(osc.ids[b], tf7[p, j, :4], tf7[p, j, 4:]); eng.set_body_transform(sb.body_ids[b], tf7[p, j, :4], tf7[p, j, 4:])
    w.set_transform(osc.ids[obj], *sc.body_tran[obj]); eng.set_body_transform(sb.body_ids[obj], *sc.body_tran[obj])
    for j, b in enumerate(order):
        tl = T.make(tf7[p, j, :4][None], tf7[p, j, 4:][None])
        dr, p1r, p2r = w.body_distance(osc.ids[b], osc.ids[obj])
        mp1r, mp2r = T.apply(tl, p1r[None])[0], T.apply(to, p2r[None])[0]
        de, l1, l2 = eng.body_to_body_distance(sb.body_ids[b], sb.body_ids[obj], local=True)
        dp = max(np.linalg.norm(l1 - mp1r), np.linalg.norm(l2 - mp2r))
        if dp > 1e-3 or abs(de - dr) > 1e-9 * max(1.0, dr):
            db, b1, b2, tp = w.brute_distance(osc.ids[b], osc.ids[obj])
            # the exhaustive winner as a two-triangle world: does the engine's own leaf arithmetic agree with the reference's?
            ta = eng.add_body(sc.bodies[b].tris[tp[0]:tp[0] + 1]); tb_ = eng.add_body(sc.bodies[obj].tris[tp[1]:tp[1] + 1])
            eng.set_body_transform(ta, tf7[p, j, :4], tf7[p, j, 4:]); eng.set_body_transform(tb_, *sc.body_tran[obj])
            dpair, q1, q2 = eng.body_to_body_distance(ta, tb_, local=True)
            print("    engine on the exhaustive winner alone: %.10f (|pts - brute pts| %.3g)" % (dpair, max(np.linalg.norm(q1 - b1), np.linalg.norm(q2 - b2))))
            print("  pose %d link %d: ref %.10f engine %.10f brute %.10f | (ref-brute)/brute %.3g (engine-brute)/brute %.3g | batch link_dist %.8f | |engine pts - brute pts| %.4g, |ref pts - brute pts| %.4g, brute tris %s"
                  % (p, j, dr, de, db, (dr - db) / db, (de - db) / db, live["link_dist"][p, j],
                     max(np.linalg.norm(l1 - b1), np.linalg.norm(l2 - b2)), max(np.linalg.norm(mp1r - b1), np.linalg.norm(mp2r - b2)), tp))
