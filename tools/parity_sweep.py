#!/usr/bin/env python
"""Wide parity sweep on the GPU box: N random planner states against the reference (oracle/_ref), PRESET and LIVE.
usage: python tools/parity_sweep.py [n_preset] [n_live] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graspit_b200.engine import B2C_INPUT_AA_EIGEN, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import fk as ofk
from tests.common import OracleScene, object_body, sample_aa_states

npre = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nlive = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 31337
sc = load_pack("barrett_mug"); obj = object_body(sc); r = sc.robots[0]
eng = Engine(0); sb = SceneBinding(eng, sc)
osc = OracleScene(sc, nthreads=ref.hardware_threads())
rng = np.random.default_rng(seed)
pos, _ = sample_aa_states(sc, npre, seed=seed, strata=("far", "near", "close"))
states = np.concatenate([pos, rng.uniform(-4, 4, size=(npre, r.eg.shape[0]))], 1).astype(np.float32)
s64 = states.astype(np.float64)
base = ofk.aa_state_to_base(r, sc.body_tran[obj], s64[:, :6])
raw64 = np.concatenate([base[0], base[1], ofk.eigen_to_dof(r, s64[:, 6:])], 1)
rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-9)
t = time.time(); lo, eo, _, _, _ = osc.eval(0, obj, raw64, mode=0); print("oracle preset %.1fs" % (time.time() - t))
res = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], states, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[obj])
m = (lo == 1) & (res["legal"] == 1)
e = rel(res["energy"][m], eo[m])
print("PRESET n=%d legal mismatches=%d  legal=%d  energy rel err: max %.3g  >1e-4: %d  >1e-5: %d" % (npre, int((lo != res["legal"]).sum()), int(m.sum()), e.max(), int((e > 1e-4).sum()), int((e > 1e-5).sum())))
t = time.time(); ll, el, dl, _, _ = osc.eval(0, obj, raw64[:nlive], mode=1, want_dist=True); print("oracle live %.1fs" % (time.time() - t))
live = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], states[:nlive], input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[obj], live=True, link_dist=True)
m = (ll == 1) & (live["legal"] == 1)
e = rel(live["energy"][m], el[m])
ld = live["link_dist"].astype(np.float64); both = (dl >= 0) & (ld >= 0)
d = rel(ld[both], dl[both])
dabs = np.abs(ld[both] - dl[both])
# relative error is meaningless for micro-distances (a 1e-9 relative difference of the link pose is 4e-7 mm): report those apart
big = dl[both] >= 0.05
print("LIVE n=%d legal mismatches=%d  energy rel err: max %.3g  >1e-4: %d | link dist sign mismatches=%d rel err max %.3g >1e-4: %d | d >= 0.05 mm: rel err max %.3g; d < 0.05 mm: abs err max %.3g mm" % (
    nlive, int((ll != live["legal"]).sum()), e.max(), int((e > 1e-4).sum()), int(((dl < 0) != (ld < 0)).sum()), d.max(), int((d > 1e-4).sum()),
    d[big].max(), dabs[~big].max() if (~big).any() else 0.0))
# single-pose pointToBodyDistance: closest points (not only distances) against the reference
w = osc.w
w.set_transform(osc.ids[obj], *sc.body_tran[obj]); eng.set_body_transform(sb.body_ids[obj], *sc.body_tran[obj])
cen = sc.bodies[obj].tris.reshape(-1, 3).mean(0) + sc.body_tran[obj][1]
npt = 20000
v = rng.normal(size=(npt, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
pts = cen[None, :] + v * rng.uniform(5, 300, size=(npt, 1))
bad = worse = 0; maxrel = 0.0
for pnt in pts:
    do, cpo, _ = w.point_distance(osc.ids[obj], pnt)
    de, cpe, _ = eng.point_to_body_distance(sb.body_ids[obj], pnt)
    maxrel = max(maxrel, abs(de - do) / max(do, 1e-9))
    if np.linalg.norm(cpe - cpo) > 1e-3:
        bad += 1
        if de > do * (1 + 1e-12):
            worse += 1
print("POINT n=%d dist rel err max %.3g | closest point differs > 1e-3 mm: %d (of which engine distance larger than the reference's: %d)" % (npt, maxrel, bad, worse))
