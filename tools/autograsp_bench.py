#!/usr/bin/env python
"""Throughput of the batched autograsp (b2c_autograsp_batch) on Barrett-vs-mug: poses/s, host-loop rounds.
usage: python tools/autograsp_bench.py [npose ...]   (GPU box)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graspit_b200.engine import AG_INTERP_ERROR, Engine, SceneBinding          # noqa: E402
from graspit_b200.formats.scenepack import load_pack                            # noqa: E402
from tests.common import aa_to_raw_states, object_body, sample_aa_states        # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 65536]
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    rid, obj = sb.robot_ids[0], sb.body_ids[object_body(sc)]
    for n in sizes:
        pos, dofs = sample_aa_states(sc, 3 * n, seed=7, strata=("near",))
        dofs[:, 1:] = 0.0
        raw = aa_to_raw_states(sc, pos, dofs)
        raw = raw[eng.eval_batch(rid, obj, raw)["legal"] == 1][:n]
        eng.auto_grasp(rid, sc.robots[0], raw)                # warm-up at full size: scratch buffers allocated outside the timed call
        l0 = eng.launch_count
        t = time.perf_counter()
        out = eng.auto_grasp(rid, sc.robots[0], raw)
        dt = time.perf_counter() - t
        print(json.dumps({"metric": "autograsps/s (Hand::autoGrasp to contact, Barrett vs mug, host buffers)", "npose": len(raw),
                          "seconds": dt, "value": len(raw) / dt, "rounds": out["rounds"], "launches": eng.launch_count - l0,
                          "mean_steps": float(out["iters"].mean()), "errors": int((out["status"] & AG_INTERP_ERROR != 0).sum()),
                          "fingers_stopped_mean": float((out["stopped"] != 0).any(axis=1).mean())}))


if __name__ == "__main__":
    main()
