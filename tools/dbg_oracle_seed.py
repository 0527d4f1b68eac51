"""How much distance-kernel work is unavoidable with the current bounds?  Variant built with -DDIST_ORACLE_SEED starts
every query from the previous call's answer: the second call's counters are the 'necessary' work."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from graspit_b200.engine import Engine, SceneBinding, B2C_INPUT_AA_EIGEN
from graspit_b200.formats.scenepack import load_pack
from graspit_b200.planner import SearchVariables
from tests.common import object_body
from bench import start_states
scene = load_pack("barrett_mug"); obj = object_body(scene)
eng = Engine(0); sb = SceneBinding(eng, scene)
sv = SearchVariables.axis_angle_eigen(2)
st = start_states(sv, 65536, 1234)
args = dict(input_mode=B2C_INPUT_AA_EIGEN, ref_tran=scene.body_tran[obj])
eng.set_profiling(True)
for i in range(3):
    r = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], st, live=True, **args)
    s = eng.last_stats(); ms = eng.last_kernel_ms()
    print("call", i, "legal", int(r["legal"].sum()), "dbox", s["dist_box_tests"], "dtri", s["dist_tri_tests"], "dist ms", round(ms["dist"], 2))
