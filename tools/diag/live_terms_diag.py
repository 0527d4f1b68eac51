"""Diagnostic (GPU box): for the LIVE outliers of the 60k sweep, the per-link energy terms of both sides."""
import os, sys
import numpy as np
np.set_printoptions(precision=7, suppress=True, linewidth=220)
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
rel = np.abs(live["energy"] - el) / np.maximum(np.abs(el), 1e-9)
bad = np.nonzero((ll == 1) & (rel > 1e-4))[0]
print("outliers", bad.tolist(), rel[bad])
w = osc.w; o_id, o_e = osc.ids[obj], sb.body_ids[obj]
to = T.make(sc.body_tran[obj][0][None], sc.body_tran[obj][1][None])
for p in bad:
    for j, b in enumerate(order):
        w.set_transform(osc.ids[b], tf7[p, j, :4], tf7[p, j, 4:]); eng.set_body_transform(sb.body_ids[b], tf7[p, j, :4], tf7[p, j, 4:])
    w.set_transform(o_id, *sc.body_tran[obj]); eng.set_body_transform(o_e, *sc.body_tran[obj])
    print("pose", p, "E engine %.7f ref %.7f" % (live["energy"][p], el[p]))
    te = []; tr = []
    for j, b in enumerate(order):
        tl = T.make(tf7[p, j, :4][None], tf7[p, j, 4:][None])
        dr, p1r, p2r = w.body_distance(osc.ids[b], o_id)
        m1, m2 = T.apply(tl, p1r[None])[0], T.apply(to, p2r[None])[0]
        de, l1, l2 = eng.body_to_body_distance(sb.body_ids[b], o_e, local=True)
        out = []
        for (a1, a2, side) in ((m1, m2, "ref"), (l1, l2, "eng")):
            n = (a1 - a2) / np.linalg.norm(a1 - a2)
            if side == "ref":
                d, cp, _ = w.point_distance(o_id, a1)
            else:
                d, cp, _ = eng.point_to_body_distance(o_e, a1)
            v = cp - a1
            out.append((np.linalg.norm(v) + 50.0 * (1.0 - np.dot(-n, v) / np.linalg.norm(v)), cp, d))
        # cross: reference closest-point query from the ENGINE's location
        dx, cpx, _ = w.point_distance(o_id, l1)
        tr.append(out[0][0]); te.append(out[1][0])
        flag = " <--" if abs(out[0][0] - out[1][0]) > 1e-4 * abs(out[0][0]) else ""
        print("  link %d term ref %.7f eng %.7f | |loc diff| %.2e | cp ref %s eng %s | ref query from engine loc: cp %s%s" % (j, out[0][0], out[1][0], np.linalg.norm(m1 - l1), out[0][1], out[1][1], cpx, flag))
    print("  mean ref %.7f eng(single-pose pieces) %.7f" % (np.mean(tr), np.mean(te)))
