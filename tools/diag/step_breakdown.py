"""Diagnostic (GPU box): where does a planner step spend its time?  CUDA-event time of N steps for growing subsets of the
step (anneal only / + publish / + collect / + host merge), and host wall time per call."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graspit_b200.engine import Annealer, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from graspit_b200.planner import BEST_LIST_SIZE, BestList, PeerExchange, SearchVariables, SimAnnParams
from graspit_b200.sampling import object_body

sc = load_pack("barrett_mug"); obj = object_body(sc)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
eng = Engine(0); sb = SceneBinding(eng, sc); eng.set_stream(stream.cuda_stream)
sv = SearchVariables.axis_angle_eigen(sc.robots[0].eg.shape[0])
rng = np.random.default_rng(1234)
init = rng.uniform(sv.vmin, sv.vmax, size=(65536, 8)).astype(np.float32)
an = Annealer(eng, sb.robot_ids[0], sb.body_ids[obj], sv, SimAnnParams.ANNEAL_DEFAULT(), init, sc.body_tran[obj], seed=1234)
px = PeerExchange(eng, an, 0, 1, device=dev)
bl = BestList(sc.body_tran[obj])
host = torch.zeros((1, BEST_LIST_SIZE, 9), dtype=torch.float32).pin_memory()
for _ in range(10):
    an.step(1); an.publish()
torch.cuda.synchronize()

def run(name, fn, n=30):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    wall = []
    for i in range(n):
        torch.cuda.synchronize()
        ev[i][0].record(stream)
        t0 = time.perf_counter()
        fn()
        wall.append(time.perf_counter() - t0)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    print("%-34s gpu-event %.4f ms/step   host time inside %.4f ms" % (name, float(np.mean(ms)), 1e3 * float(np.mean(wall))))

run("anneal step", lambda: an.step(1))
run("anneal step + publish", lambda: (an.step(1), an.publish()))
run("+ collect (async)", lambda: (an.step(1), an.publish(), eng.exchange_collect(host.data_ptr(), 0, True)))
def full():
    an.step(1); an.publish(); eng.exchange_collect(host.data_ptr(), 0, True)
    r = host.numpy().reshape(-1, 9); e = r[:, -1]; m = e < min(bl.worst_energy(), 1e7)
    if m.any(): bl.merge(r[m, :-1], e[m])
run("+ host merge", full)
run("publish alone", lambda: an.publish())
run("collect alone (sync)", lambda: eng.exchange_collect(host.data_ptr(), 0, False))
