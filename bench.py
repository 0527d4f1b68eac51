#!/usr/bin/env python
"""Headline benchmark: hand-pose evaluations / second (FK + collision legality + CONTACT_ENERGY).

Workload (BASELINE.json configs[1]): Barrett hand vs mug (worlds/plannerMug.xml), the
EGPlanner's SimAnnPlanner with CONTACT_ENERGY: 65 536 independent annealing chains per GPU in the planner dialog's
default state layout (SPACE_AXIS_ANGLE position + POSE_EIGEN posture = 8 fp32 variables), SimAnnParams
ANNEAL_DEFAULT.  One step = one annealing step of every chain = one pass of the hot path over one batch:
    neighbour generation (Philox-keyed Ingber distribution) -> FK -> legality over the 35 active pairs ->
    [CONTACT_LIVE: 9 body-body distances] -> CONTACT_ENERGY -> Metropolis accept,
all on the device with the chain states resident in HBM (`value`).  `e2e` is the same batches evaluated through the
C ABI with HOST buffers (b2c_eval_batch: pinned H2D of the neighbour states, kernels, D2H of legal + energy) -- the
call a GraspIt!-side batched SearchEnergy makes.  Chains start uniformly inside the planner's search box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode preset|live] [--impl reference] [--workload anneal|random]

N > 1 is launched by torchrun (one rank per GPU); chains shard across ranks with no data-path collective, and each
step ends with the small NCCL all-gather of every rank's best (state || energy) rows that the planner's best-list
merge needs.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hand-pose evals/sec (collision+contact+energy)"
UNIT = "poses/s"
BEST_LIST_SIZE = 20          # SimAnnPlanner BEST_LIST_SIZE, reference src/EGPlanner/simAnnPlanner.cpp:36
ALGO_BYTES_EVAL_PER_POSE = 9 * 96 + 1 + 4     # eval kernel: 9 link transforms (12 fp64) in, legal + energy out
ALGO_BYTES_STEP_PER_POSE = 8 * 4 + 1 + 4 + 2 * 9 * 96   # whole step: states in, legal+energy out, transforms written+read


def sample_states(scene, n, seed):
    """(n, 8) float32 planner states: Tx Ty Tz theta phi alpha + 2 eigengrasp amplitudes."""
    from tests.common import sample_aa_states
    pos, _ = sample_aa_states(scene, n, seed=seed)
    rng = np.random.default_rng(seed + 7919)
    neg = scene.robots[0].eg.shape[0]
    amps = rng.uniform(-4.0, 4.0, size=(n, neg)).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([pos, amps], 1).astype(np.float32))


def states_to_raw(scene, states):
    """oracle-side state -> (base, dofs) mapping (fp64 on the widened fp32 states)"""
    from oracle.port import fk as ofk
    from tests.common import object_body
    r = scene.robots[0]
    s = states.astype(np.float64)
    base = ofk.aa_state_to_base(r, scene.body_tran[object_body(scene)], s[:, :6])
    dofs = ofk.eigen_to_dof(r, s[:, 6:])
    return np.concatenate([base[0], base[1], dofs], 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def start_states(sv, n, seed):
    """chains start uniformly inside the planner's search box (SearchVariable ranges)"""
    rng = np.random.default_rng(seed)
    return rng.uniform(sv.vmin, sv.vmax, size=(n, len(sv.vmin))).astype(np.float32)


def cpu_neighbour_batch(sv, params, n, step, seed, chain_offset=0):
    """a batch of annealing neighbours made on the CPU (oracle port of SimAnn::variableNeighbor) from the start states:
    the same distribution the device annealer evaluates in its first steps"""
    from oracle.port import simann
    cur = start_states(sv, n, seed)
    return simann.propose(cur, np.full(n, params.K0), sv, params, step=step, seed=seed, chain_offset=chain_offset)


class CpuReference:
    """the reference's own CPU implementation of the path (oracle/_ref: GraspitCollision compiled verbatim, one
    instance per host thread) evaluating planner states; FK is precomputed outside the timed call"""

    def __init__(self, scene, threads):
        from tests.common import OracleScene, object_body
        self.scene, self.osc, self.obj = scene, OracleScene(scene, nthreads=threads), object_body(scene)

    def time(self, states, mode):
        from oracle import ref
        raw = states_to_raw(self.scene, states)
        order, tf7 = self.osc.link_poses(0, raw)
        r = self.scene.robots[0]
        hand_ids = [self.osc.ids[b] for b in order]
        vc_link = [order.index(v.body) for v in r.vcontacts]
        vc_loc = np.array([v.loc for v in r.vcontacts]).reshape(-1, 3)
        vc_n = np.array([v.normal for v in r.vcontacts]).reshape(-1, 3)
        t0 = time.perf_counter()
        ref.eval_batch(self.osc.worlds, hand_ids, self.osc.ids[self.obj], tf7, vc_link, vc_loc, vc_n, mode=mode)
        return time.perf_counter() - t0


def cpu_sample_size(args):
    return args.cpu_sample if args.mode != "live" else max(256, args.cpu_sample // 16)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from graspit_b200.formats.scenepack import load_pack
    from graspit_b200.planner import SearchVariables, SimAnnParams
    from oracle import ref
    scene = load_pack("barrett_mug")
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgraspit_ref.so missing"}))
        return
    threads = max(1, ref.hardware_threads())
    mode = 1 if args.mode == "live" else 0
    sample = cpu_sample_size(args)
    sv = SearchVariables.axis_angle_eigen(scene.robots[0].eg.shape[0])
    params = SimAnnParams.ANNEAL_DEFAULT()
    cpu = CpuReference(scene, threads)
    times = []
    for step in range(args.warmup + args.steps):
        st = cpu_neighbour_batch(sv, params, sample, step, 1234) if args.workload == "anneal" else sample_states(scene, sample, 1234 + step)
        dt = cpu.time(st, mode)
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.poses),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{sample} neighbour states per step of the same workload (reference GraspitCollision compiled "
                                       "verbatim, -O3, one instance per thread; FK precomputed outside the timed call)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, poses):
    mode = "CONTACT_LIVE" if args.mode == "live" else "CONTACT_PRESET"
    if args.workload == "anneal":
        w = ("Barrett hand vs mug (worlds/plannerMug.xml), EGPlanner SimAnnPlanner CONTACT_ENERGY " + mode +
             f", {poses} annealing chains per GPU, one neighbour per chain per step (SimAnnParams ANNEAL_DEFAULT, "
             "SPACE_AXIS_ANGLE + POSE_EIGEN = 8 fp32 variables, Philox seed 1234, chains started uniformly in the planner's box)")
    else:
        w = ("Barrett hand vs mug (worlds/plannerMug.xml), CONTACT_ENERGY " + mode + f", {poses} random planner states per step per GPU "
             "(SPACE_AXIS_ANGLE + POSE_EIGEN, 8 fp32 variables; 50% planner-box / 50% within 150 mm strata)")
    return {"workload": w, "poses_per_step_per_gpu": poses, "mode": args.mode, "kind": args.workload,
            "l2": "flushed between timed steps (256 MiB write); per-step inputs 2 MiB"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="preset", choices=["preset", "live"])
    ap.add_argument("--workload", default="anneal", choices=["anneal", "random"])
    ap.add_argument("--poses", type=int, default=65536)
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from graspit_b200.engine import B2C_EVAL_LIVE, B2C_INPUT_AA_EIGEN, Annealer, Engine, EvalArgs, SceneBinding
    from graspit_b200.formats.scenepack import load_pack
    from graspit_b200.planner import BEST_LIST_SIZE, SearchVariables, SimAnnParams, exchange_best, local_best_rows
    from tests.common import object_body
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scene = load_pack("barrett_mug")
    obj = object_body(scene)
    eng = Engine(local_rank)
    sb = SceneBinding(eng, scene)
    rid, oid = sb.robot_ids[0], sb.body_ids[obj]
    ref_tran = scene.body_tran[obj]
    live = args.mode == "live"
    npose, K, W = args.poses, args.steps, args.warmup
    nvar = 6 + scene.robots[0].eg.shape[0]
    sv = SearchVariables.axis_angle_eigen(scene.robots[0].eg.shape[0])
    params = SimAnnParams.ANNEAL_DEFAULT()

    # one explicit stream for torch ops, the engine's kernels and the timing events
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    anneal = args.workload == "anneal"
    an = None
    if anneal:
        # chains [rank * npose, (rank + 1) * npose) of world * npose: weak scaling, chains never migrate
        an = Annealer(eng, rid, oid, sv, params, start_states(sv, npose, 1234 + rank), ref_tran, seed=1234,
                      chain_offset=rank * npose, live=live)
        views = an.torch_views(dev)
        # host copies of the neighbour batches of K + W consecutive annealing steps, for the end-to-end leg
        host_batches = []
        for _ in range(K + W):
            host_batches.append(an.propose())
            an.step(1)
    else:
        host_batches = [sample_states(scene, npose, 1234 + 1000 * rank + s) for s in range(K + W)]
    pinned = [torch.from_numpy(b).pin_memory() for b in host_batches]
    dev_batches = None if anneal else [p.to(dev, non_blocking=True) for p in pinned]
    legal_d = torch.zeros(npose, dtype=torch.uint8, device=dev)
    energy_d = torch.zeros(npose, dtype=torch.float32, device=dev)
    legal_h = torch.zeros(npose, dtype=torch.uint8).pin_memory()
    energy_h = torch.zeros(npose, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    eng.set_profiling(True)

    # N > 1: every step ends with the all-gather of each rank's BEST_LIST_SIZE best (state || energy) rows.  The rows
    # come from the engine's incremental best list (one-warp kernel over the chains that improved this step), so the
    # exchange costs one tiny kernel + one 720-byte-per-rank NCCL all-gather
    # The exchange of step s overlaps the kernels of step s + 1: the all-gather is asynchronous, and a rank
    # only waits for it -- on the stream, not on the host -- after it has enqueued its next step.  Rows and results are
    # kept in a ring of three; the merged list trails the chains by up to two steps, which the planner's best-list rule does not mind.
    DEPTH = 3                                   # exchanges in flight: a rank never waits for a gather younger than two steps
    rows = [torch.empty((BEST_LIST_SIZE, nvar + 1), dtype=torch.float32, device=dev) for _ in range(DEPTH)] if anneal else None
    gathered = [torch.empty((world * BEST_LIST_SIZE, nvar + 1), dtype=torch.float32, device=dev) for _ in range(DEPTH)] if world > 1 else None
    pending = []

    def step_device(i):
        if anneal:
            an.step(1)
            if world > 1:
                if len(pending) >= DEPTH - 1:
                    pending.pop(0).wait()
                an.best_rows(BEST_LIST_SIZE, rows[i % DEPTH].data_ptr())
                pending.append(dist.all_gather_into_tensor(gathered[i % DEPTH], rows[i % DEPTH], async_op=True))
        else:
            st = dev_batches[i]
            eng.eval_batch_device(rid, oid, st.data_ptr(), npose, nvar, legal_d.data_ptr(), energy_d.data_ptr(),
                                  input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ref_tran, live=live)
            if world > 1:
                e = torch.where(legal_d.bool(), energy_d, torch.full_like(energy_d, 3.0e38))
                exchange_best(local_best_rows(st, e, BEST_LIST_SIZE))

    def sync_all():
        while pending:
            pending.pop(0).wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step_device(i)
    sync_all()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kernel_ms = []
    sync_all()
    for s in range(K):
        flush.fill_(s & 0xff)                      # L2 flush, outside the event bracket
        ev[s][0].record(stream)
        step_device(W + s)
        ev[s][1].record(stream)
        kernel_ms.append(eng.last_kernel_ms())     # synchronises; after the bracket
    sync_all()
    launches = eng.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    stats = eng.last_stats()
    anneal_info = None
    if anneal:
        d = an.read()
        anneal_info = {"steps_taken": d["steps"], "neighbours_evaluated": d["evaluated"], "legal_fraction": d["legal_neighbours"] / max(1, d["evaluated"]),
                       "jump_fraction": d["jumps"] / max(1, d["evaluated"]),
                       "best_energy_min": float(d["best_energy"].min()), "best_energy_median": float(np.median(d["best_energy"]))}

    # end to end through the C ABI with host buffers (pinned), H2D + kernels + D2H inside the timed region
    eng.set_profiling(False)

    def step_host(i):
        a = EvalArgs()
        a.robot, a.object, a.npose, a.input_mode = rid, oid, npose, B2C_INPUT_AA_EIGEN
        a.states, a.stride = pinned[i].data_ptr(), nvar
        a.ref_q[:] = list(ref_tran[0]); a.ref_t[:] = list(ref_tran[1])
        a.flags = B2C_EVAL_LIVE if live else 0
        a.legal, a.energy = legal_h.data_ptr(), energy_h.data_ptr()
        eng._ck(eng.L.b2c_eval_batch(eng.h, C.byref(a)))

    for i in range(W):
        step_host(i)
    sync_all()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for s in range(K):
        flush.fill_(s & 0xff)
        ev2[s][0].record(stream)
        step_host(W + s)
        ev2[s][1].record(stream)
    sync_all()
    e2e_ms = float(sum(a.elapsed_time(b) for a, b in ev2))
    nlegal = int(legal_h.sum().item())
    clocks = sampler.stop()

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / K
    value = world * npose / (ms_per_step / 1e3)
    e2e_value = world * npose / (e2e_ms / K / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        kms = {k: float(np.mean([x[k] for x in kernel_ms])) for k in kernel_ms[0]}
        dom = max(kms, key=kms.get)
        dom_ms = kms[dom]
        # algorithmic DRAM bytes per launch: what the kernel must read and write once
        algo = {"eval": (9 * 96 + 1 + 4) * npose,                 # 9 link transforms in, legal + energy(0) out
                "energy": nlegal * (9 * 96 + 4 + 4) + npose,      # legal list + transforms in, energy out
                "dist": nlegal * 9 * (2 * 96 + 4 + 48),           # per (legal pose, link): 2 transforms in, dist + 2 points out
                "fk": npose * (nvar * 4 + 9 * 96), "recheck": 0}
        algo_bytes = algo.get(dom, 0)
        achieved = algo_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(args.mode, {}).get(f"{dom}_kernel")
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args, npose),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": npose * nvar * 4,
                        "d2h_bytes_per_step": npose * 5, "ms_per_step": e2e_ms / K},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": algo_bytes,
                             "kernel_ms_all": kms, "legal_fraction_last_step": nlegal / npose,
                             "note": "working set is L2-resident by design: compulsory HBM traffic only (SURVEY.md 8(d)); the kernel is "
                                     "bound by instruction issue under divergence, see profiles/; traffic = dram bytes of one ncu "
                                     "capture (profiles/traffic.json); traversal work of the last step: " + json.dumps(stats)},
                }
        if anneal_info:
            line["anneal"] = anneal_info
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref
                threads = max(1, ref.hardware_threads())
                sample = cpu_sample_size(args)
                cpu = CpuReference(scene, threads)
                cpu.time(host_batches[W][: max(64, sample // 8)], 1 if live else 0)            # warm-up
                secs = min(cpu.time(host_batches[W][:sample], 1 if live else 0) for _ in range(3))
                line["cpu_baseline"] = {"value": sample / secs, "unit": UNIT, "cores": threads, "kind": "reference",
                                        "sample": f"first {sample} states of the first timed batch, best of 3 ({secs:.2f} s); reference "
                                                  "GraspitCollision compiled verbatim (-O3), one instance per thread, FK precomputed"}
            except Exception as ex:   # the oracle is a checker, never a dependency of the measurement
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
