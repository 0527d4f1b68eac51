"""Diagnostic (GPU box): where does one end-to-end b2c_eval_batch call (host buffers) spend its time, against the same
batch through device pointers?"""
import os, sys, time
import ctypes as C
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graspit_b200.engine import B2C_INPUT_AA_EIGEN, Engine, EvalArgs, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from graspit_b200.planner import SearchVariables
from graspit_b200.sampling import object_body

sc = load_pack("barrett_mug"); obj = object_body(sc)
eng = Engine(0); sb = SceneBinding(eng, sc)
rid, oid = sb.robot_ids[0], sb.body_ids[obj]
sv = SearchVariables.axis_angle_eigen(sc.robots[0].eg.shape[0])
rng = np.random.default_rng(1234)
n = 65536
st = torch.from_numpy(rng.uniform(sv.vmin, sv.vmax, size=(n, 8)).astype(np.float32)).pin_memory()
lg = torch.zeros(n, dtype=torch.uint8).pin_memory(); en = torch.zeros(n, dtype=torch.float32).pin_memory()
ref = sc.body_tran[obj]
a = EvalArgs()
a.robot, a.object, a.npose, a.input_mode = rid, oid, n, B2C_INPUT_AA_EIGEN
a.states, a.stride = st.data_ptr(), 8
a.ref_q[:] = list(ref[0]); a.ref_t[:] = list(ref[1])
a.legal, a.energy = lg.data_ptr(), en.data_ptr()
for _ in range(5):
    eng._ck(eng.L.b2c_eval_batch(eng.h, C.byref(a)))
eng.set_profiling(True)
w = []
for _ in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng._ck(eng.L.b2c_eval_batch(eng.h, C.byref(a)))
    w.append(time.perf_counter() - t0)
k = eng.last_kernel_ms()
print("host-buffer call: wall %.4f ms (min %.4f)   kernels of the last call: %s  sum %.4f" % (1e3 * np.mean(w), 1e3 * min(w), {x: round(y, 4) for x, y in k.items()}, sum(k.values())))
eng.set_profiling(False)
w = []
for _ in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng._ck(eng.L.b2c_eval_batch(eng.h, C.byref(a)))
    w.append(time.perf_counter() - t0)
print("host-buffer call, profiling off: wall %.4f ms (min %.4f)" % (1e3 * np.mean(w), 1e3 * min(w)))
