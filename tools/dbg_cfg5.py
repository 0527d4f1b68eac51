import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from graspit_b200.engine import Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from tests.common import *
small, big = load_pack("barrett_mug"), load_pack("barrett_mug_200k")
e1, e2 = Engine(0), Engine(0)
s1, s2 = SceneBinding(e1, small), SceneBinding(e2, big)
obj = object_body(small)
pos, dofs = sample_aa_states(small, 20000, seed=4321)
raw = aa_to_raw_states(small, pos, dofs)
r1 = e1.eval_batch(s1.robot_ids[0], s1.body_ids[obj], raw, link_dist=True)
r2 = e2.eval_batch(s2.robot_ids[0], s2.body_ids[obj], raw, link_dist=True)
d1, d2 = r1["link_dist"], r2["link_dist"]
rel = np.abs(d1 - d2) / np.maximum(np.abs(d1), 1e-9)
bad = np.argwhere((rel > 1e-4) & (d1 > 1e-2) & (d2 > 1e-2))
print("mismatches", len(bad), "of", d1.size)
poses = sorted(set(int(b[0]) for b in bad))[:40]
osc = OracleScene(small, nthreads=4)
_, _, dist_o, _, _ = osc.eval(0, obj, raw[poses], mode=0, want_dist=True)
for i, p in enumerate(poses):
    for l in range(d1.shape[1]):
        if rel[p, l] > 1e-4:
            print(p, l, "small", d1[p, l], "big", d2[p, l], "oracle(small)", dist_o[i, l])
