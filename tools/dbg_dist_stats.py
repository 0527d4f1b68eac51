"""Debug-variant counters of the distance kernel (build with -DDIST_DEBUG_STATS, run with B2C_LIB=that build)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from graspit_b200.engine import Engine, SceneBinding, B2C_INPUT_AA_EIGEN
from graspit_b200.formats.scenepack import load_pack
from tests.common import object_body
from bench import sample_states
scene = load_pack("barrett_mug"); obj = object_body(scene)
eng = Engine(0); sb = SceneBinding(eng, scene)
st = sample_states(scene, 65536, 1234)
args = dict(input_mode=B2C_INPUT_AA_EIGEN, ref_tran=scene.body_tran[obj])
eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st, live=False, **args)
s0 = eng.last_stats()
res = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st, live=True, **args)
s1 = eng.last_stats()
print("preset", s0); print("live", s1)
nl = int(res["legal"].sum())
print("legal poses %d -> %d queries" % (nl, nl * 9))
print("early_out=%d improved=%d beyond10pct=%d while_best>50mm=%d of dtri=%d dbox=%d" % (
    s1["box_tests"] - s0["box_tests"], s1["tri_tests"] - s0["tri_tests"], s1["ambiguous"] - s0["ambiguous"],
    s1["recheck_hits"] - s0["recheck_hits"], s1["dist_tri_tests"], s1["dist_box_tests"]))
# cost vs distance: histogram of the returned link distances
r2 = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st[:4096], link_dist=True, **args)
d = r2["link_dist"]; print("link distance quantiles (mm):", np.quantile(d[d >= 0], [0.1, 0.25, 0.5, 0.75, 0.9]))
