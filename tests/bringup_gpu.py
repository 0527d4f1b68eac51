"""First-light script for the GPU box: engine vs oracle on the config-1 pose batch (not a pytest)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graspit_b200.formats.scenepack import load_pack
from graspit_b200.engine import Engine, SceneBinding, B2C_INPUT_RAW
from tests.common import *

sc = load_pack("barrett_mug")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pos, dofs = sample_aa_states(sc, n)
raw = aa_to_raw_states(sc, pos, dofs)
osc = OracleScene(sc, nthreads=8)
obj = object_body(sc)
t0 = time.time(); legal_o, energy_o, dist_o, order, tf7 = osc.eval(0, obj, raw, mode=0, want_dist=True); t1 = time.time()
print(f"oracle: {n} poses in {t1-t0:.2f}s legal={legal_o.sum()}")
eng = Engine(0)
sb = SceneBinding(eng, sc)
rid = sb.robot_ids[0]
print("robot bodies", eng.robot_bodies(rid), "pairs", len(eng.robot_pairs(rid)))
t0 = time.time()
res = eng.eval_batch(rid, sb.body_ids[obj], raw, pair_mask=True, link_dist=True, body_tf=True)
t1 = time.time()
print(f"engine: {t1-t0:.3f}s legal={res['legal'].sum()} stats={eng.last_stats()}")
# FK parity
btf = res["body_tf"]
err_t = np.abs(btf[:, :, 4:7] - tf7[:, :, 4:7]).max()
qd = np.minimum(np.abs(btf[:, :, 0:4] - tf7[:, :, 0:4]).max(), np.abs(btf[:, :, 0:4] + tf7[:, :, 0:4]).max())
print("FK max |dt|", err_t, "max |dq|", qd)
print("legal mismatches:", int((res["legal"] != legal_o).sum()))
m = (legal_o == 1) & (res["legal"] == 1)
rel = np.abs(res["energy"][m] - energy_o[m]) / np.maximum(np.abs(energy_o[m]), 1e-9)
print("energy max rel err:", rel.max() if m.any() else None, "mean", rel.mean() if m.any() else None)
ld = res["link_dist"]
both = (dist_o >= 0) & (ld >= 0)
print("dist sign mismatches:", int(((dist_o < 0) != (ld < 0)).sum()))
rd = np.abs(ld[both] - dist_o[both]) / np.maximum(dist_o[both], 1e-9)
print("dist max rel err:", rd.max(), "mean", rd.mean())
# pair sets
ps = osc.pair_sets(0, raw[:200])
pairs = eng.robot_pairs(rid)
inv = {v: k for k, v in sb.body_ids.items()}
bad = 0
for p in range(200):
    got = set()
    for i, (a, b) in enumerate(pairs):
        if (int(res["pair_mask"][p, i >> 6]) >> (i & 63)) & 1:
            got.add(tuple(sorted((inv[a], inv[b]))))
    if got != ps[p]:
        bad += 1
        if bad < 5: print("pair mismatch pose", p, got ^ ps[p])
print("pair-set mismatches (first 200):", bad)
# LIVE
res2 = eng.eval_batch(rid, sb.body_ids[obj], raw, live=True)
legal_l, energy_l, _, _, _ = osc.eval(0, obj, raw, mode=1)
m = (legal_l == 1) & (res2["legal"] == 1)
rel = np.abs(res2["energy"][m] - energy_l[m]) / np.maximum(np.abs(energy_l[m]), 1e-9)
print("LIVE energy max rel err:", rel.max(), "mean", rel.mean(), "stats", eng.last_stats())
# single-pose API
d, cp, cn = eng.point_to_body_distance(sb.body_ids[obj], [10, 11, 12])
print("point dist", d, cp, cn, "oracle", osc.w.point_distance(osc.ids[obj], [10, 11, 12]))
print("body dist", eng.body_to_body_distance(sb.body_ids[1], sb.body_ids[obj]), "oracle", osc.w.body_distance(osc.ids[1], osc.ids[obj]))
print("bvs", eng.get_bounding_volumes(sb.body_ids[obj], 0), osc.w.bounding_volumes(osc.ids[obj], 0))
for big in (16384, 65536):
    pos, dofs = sample_aa_states(sc, big, seed=7)
    rawb = aa_to_raw_states(sc, pos, dofs)
    eng.eval_batch(rid, sb.body_ids[obj], rawb)
    t0 = time.time(); r3 = eng.eval_batch(rid, sb.body_ids[obj], rawb); t1 = time.time()
    print(f"PRESET {big} poses: {t1-t0:.4f}s -> {big/(t1-t0):.0f} poses/s; legal {r3['legal'].mean():.3f}", eng.last_stats())
    t0 = time.time(); r3 = eng.eval_batch(rid, sb.body_ids[obj], rawb, live=True); t1 = time.time()
    print(f"LIVE {big} poses: {t1-t0:.4f}s -> {big/(t1-t0):.0f} poses/s", eng.last_stats())
