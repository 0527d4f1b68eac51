"""dump distance outliers (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graspit_b200.engine import Engine, SceneBinding
from tests.common import *
sc = load_pack("barrett_mug")
pos, dofs = sample_aa_states(sc, 1000)
raw = aa_to_raw_states(sc, pos, dofs)
osc = OracleScene(sc, nthreads=8)
obj = object_body(sc)
legal_o, energy_o, dist_o, order, tf7 = osc.eval(0, obj, raw, mode=0, want_dist=True)
eng = Engine(0); sb = SceneBinding(eng, sc); rid = sb.robot_ids[0]
res = eng.eval_batch(rid, sb.body_ids[obj], raw, link_dist=True)
ld = res["link_dist"].astype(np.float64)
both = (dist_o >= 0) & (ld >= 0)
rd = np.where(both, np.abs(ld - dist_o) / np.maximum(dist_o, 1e-9), 0)
idx = np.argwhere(rd > 2e-6)
print("n outliers", len(idx))
for p, l in idx[:20]:
    # single-pose API on the same configuration
    for k, b in enumerate(order):
        eng.set_body_transform(sb.body_ids[b], tf7[p, k, :4], tf7[p, k, 4:])
        osc.w.set_transform(osc.ids[b], tf7[p, k, :4], tf7[p, k, 4:])
    d1 = eng.body_to_body_distance(sb.body_ids[order[l]], sb.body_ids[obj], local=True)
    d2 = osc.w.body_distance(osc.ids[order[l]], osc.ids[obj])
    print(p, l, "gpu batch", ld[p, l], "oracle", dist_o[p, l], "rel", rd[p, l], "single gpu", d1[0], "single oracle", d2[0])
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/dist_debug.npz", raw=raw, ld=ld, dist_o=dist_o, tf7=tf7)
