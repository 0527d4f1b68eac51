"""Diagnostic (GPU box): joint-level FK of the engine vs the restated fwdKinematics, on the closed hands of the autograsp
fixture; and the LIVE outliers of the 60k sweep with exhaustive distances."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graspit_b200.engine import AG_INTERP_ERROR, B2C_INPUT_JOINTS, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle.port.autograsp import AutoGraspOracle
from tests.common import OracleScene

sc = load_pack("barrett_mug")
eng = Engine(0); sb = SceneBinding(eng, sc); rid = sb.robot_ids[0]
raw = np.load("tests/golden/barrett_mug_autograsp.npz")["raw"]
out = eng.auto_grasp(rid, sc.robots[0], raw)
ok = (out["status"] & AG_INTERP_ERROR) == 0
js = np.concatenate([raw[ok, :7].astype(np.float64), out["joints"][ok]], axis=1)
tf = eng.eval_batch(rid, -1, js, input_mode=B2C_INPUT_JOINTS, body_tf=True)["body_tf"]
osc = OracleScene(sc, nthreads=1); ag = AutoGraspOracle(sc, 0, osc.w, osc.ids)
inv_e = {v: k for k, v in sb.body_ids.items()}
bodies = eng.robot_bodies(rid)
worst_t = np.zeros(len(bodies)); worst_q = np.zeros(len(bodies))
for p in range(len(js)):
    ag.set_joints((js[p, 0:4], js[p, 4:7]), js[p, 7:])
    for k, b in enumerate(bodies):
        qo, to = ag.last_tf[inv_e[b]]
        worst_t[k] = max(worst_t[k], np.abs(tf[p, k, 4:] - to).max())
        worst_q[k] = max(worst_q[k], min(np.abs(tf[p, k, :4] - qo).max(), np.abs(tf[p, k, :4] + qo).max()))
np.set_printoptions(linewidth=250)
print("per body max |dt|", worst_t)
print("per body max |dq|", worst_q)
print("base q norm - 1:", np.abs(np.linalg.norm(js[:, :4], axis=1) - 1).max())

# the same on the poses of the contact sweep, if present
if os.path.exists("gpurun_out/contacts_sweep_r02d.npz") or os.path.exists("tools/diag/sweep_js.npz"):
    pth = "tools/diag/sweep_js.npz" if os.path.exists("tools/diag/sweep_js.npz") else "gpurun_out/contacts_sweep_r02d.npz"
    js2 = np.load(pth)["js"][:300]
    tf2 = eng.eval_batch(rid, -1, js2, input_mode=B2C_INPUT_JOINTS, body_tf=True)["body_tf"]
    wt = np.zeros(len(bodies)); wq = np.zeros(len(bodies))
    for p in range(len(js2)):
        ag.set_joints((js2[p, 0:4], js2[p, 4:7]), js2[p, 7:])
        for k, b in enumerate(bodies):
            qo, to = ag.last_tf[inv_e[b]]
            wt[k] = max(wt[k], np.abs(tf2[p, k, 4:] - to).max())
            wq[k] = max(wq[k], min(np.abs(tf2[p, k, :4] - qo).max(), np.abs(tf2[p, k, :4] + qo).max()))
    print("sweep poses: per body max |dt|", wt)
    print("sweep poses: per body max |dq|", wq)
