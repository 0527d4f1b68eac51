#!/usr/bin/env python
"""Parity sweep of the batched autograsp against the per-pose restatement on the reference's collision backend.
  GPU box:   python tools/autograsp_sweep.py gpu N out.npz      (engine results for N seeded open-hand Barrett poses)
  anywhere:  python tools/autograsp_sweep.py cpu out.npz [procs] (oracle on the same poses, statistics as one JSON line)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graspit_b200.formats.scenepack import load_pack                            # noqa: E402
from tests.common import aa_to_raw_states, object_body, sample_aa_states        # noqa: E402


def gpu(n, out):
    from graspit_b200.engine import Engine, SceneBinding
    sc = load_pack("barrett_mug")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    rid, oid = sb.robot_ids[0], sb.body_ids[object_body(sc)]
    pos, dofs = sample_aa_states(sc, 4 * n, seed=777, strata=("near",))
    dofs = dofs.copy(); dofs[:, 1:] = 0.0
    raw = aa_to_raw_states(sc, pos, dofs)
    raw = raw[eng.eval_batch(rid, oid, raw)["legal"] == 1][:n]
    res = eng.auto_grasp(rid, sc.robots[0], raw)
    np.savez_compressed(out, raw=raw, dof=res["dof"], joints=res["joints"], status=res["status"], iters=res["iters"], stopped=res["stopped"])
    print("wrote", out, len(raw), "poses")


def _chunk(args):
    lo, hi, path = args
    from oracle.port.autograsp import AutoGraspOracle
    from tests.common import OracleScene
    z = np.load(path)
    sc = load_pack("barrett_mug")
    osc = OracleScene(sc, nthreads=1)
    ag = AutoGraspOracle(sc, 0, osc.w, osc.ids)
    rows = []
    for k in range(lo, hi):
        raw = z["raw"][k]
        r = ag.auto_grasp((raw[0:4].astype(np.float64), raw[4:7].astype(np.float64)), raw[7:].astype(np.float64))
        rows.append((k, r["error"], r["moved"], r["iters"], r["dof"], r["stopped"], r["joints"]))
    return rows


def cpu(path, procs):
    import multiprocessing as mp
    z = np.load(path)
    n = len(z["raw"])
    step = (n + 4 * procs - 1) // (4 * procs)
    jobs = [(i, min(n, i + step), path) for i in range(0, n, step)]
    with mp.Pool(procs) as pool:
        rows = [r for part in pool.map(_chunk, jobs) for r in part]
    err_flag = moved_flag = 0
    errs, same_iters, same_stopped = [], 0, 0
    for k, error, moved, iters, dof, stopped, joints in rows:
        st = int(z["status"][k])
        err_flag += bool(st & 2) != error
        moved_flag += bool(st & 1) != moved
        if error:
            continue
        e = float(np.abs(z["dof"][k] - dof).max())
        errs.append(e)
        if e < 1e-9:
            same_iters += int(z["iters"][k]) == iters
            same_stopped += bool((z["stopped"][k] == stopped).all())
    errs = np.array(errs)
    print(json.dumps({"what": "batched autograsp vs per-pose reference control flow on GraspitCollision (Barrett vs mug, open hand, near poses)",
                      "poses": n, "compared": int(len(errs)), "status_error_mismatch": err_flag, "status_moved_mismatch": moved_flag,
                      "dof_identical_1e-9": float((errs < 1e-9).mean()), "dof_within_1e-3_rad": float((errs < 1e-3).mean()),
                      "dof_max_abs_err_rad": float(errs.max()), "dof_err_p99_rad": float(np.quantile(errs, 0.99)),
                      "identical_poses_same_iterations": same_iters, "identical_poses_same_stopped_joints": same_stopped}))


if __name__ == "__main__":
    if sys.argv[1] == "gpu":
        gpu(int(sys.argv[2]), sys.argv[3])
    else:
        cpu(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else max(1, (os.cpu_count() or 2) - 1))
