#!/usr/bin/env python
"""Headline benchmark: hand-pose evaluations / second (FK + collision legality + CONTACT_ENERGY).

Workload (BASELINE.json configs[1]): Barrett hand vs mug (worlds/plannerMug.xml), the
EGPlanner's SimAnnPlanner with CONTACT_ENERGY: 65 536 independent annealing chains per GPU in the planner dialog's
default state layout (SPACE_AXIS_ANGLE position + POSE_EIGEN posture = 8 fp32 variables), SimAnnParams
ANNEAL_DEFAULT.  One step = one annealing step of every chain = one pass of the hot path over one batch:
    neighbour generation (Philox-keyed Ingber distribution) -> FK -> legality over the 35 active pairs ->
    [CONTACT_LIVE: 9 body-body distances] -> CONTACT_ENERGY -> Metropolis accept -> best-list update + exchange,
all on the device with the chain states resident in HBM (`value`).  `e2e` is the same batches evaluated through the
C ABI with HOST buffers (b2c_eval_batch: H2D of the neighbour states, kernels, D2H of legal + energy, then the best-row
exchange) -- the call a GraspIt!-side batched SearchEnergy makes.  Chains start uniformly inside the planner's search box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode preset|live] [--workload anneal|config5|random]
                    [--exchange peer|nccl] [--no-sub] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); chains shard across ranks with no data-path collective.  The planner's
exchange -- every rank's BEST_LIST_SIZE best (state || energy) rows, merged into the best list on the host -- runs every
step inside the timed region, at N = 1 too (there it is the planner's own best-list upkeep): by default through the
engine's collective-free peer exchange (rows stored straight into the peers' HBM over NVLink by the kernel that updates
the best list), `--exchange nccl` keeps the NCCL all-gather as the checked alternative.

Besides the headline line the default run adds sub-records (`--no-sub` skips them).  At N = 1 only: "config3" (allContacts
for 16 384 closed HumanHand poses, with "region_batch": bodyRegion for both sides of every contact on an elastic body), "config4" (allCollisions for 262 144 arm + hand configurations), "autograsp" (65 536
hands closed in one call) -- one timed call each, host buffers.  At every N, measured like the headline:
    "live"     CONTACT_LIVE annealing steps of the same chains,
    "config5"  BASELINE.json configs[4]: 1 048 576 chains in total (strong scaling: / N per GPU) against the mug
               subdivided to 220 800 triangles.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
CONFIG5_TOTAL = 1048576


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


# ----------------------------------------------------------------------------------------------------- reference arm
def states_to_raw(scene, states):
    """oracle-side state -> (base, dofs) mapping (fp64 on the widened fp32 states)"""
    from graspit_b200.sampling import object_body
    from oracle.port import fk as ofk
    r = scene.robots[0]
    s = states.astype(np.float64)
    base = ofk.aa_state_to_base(r, scene.body_tran[object_body(scene)], s[:, :6])
    dofs = ofk.eigen_to_dof(r, s[:, 6:])
    return np.concatenate([base[0], base[1], dofs], 1)


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
        from graspit_b200.sampling import object_body
        from tests.common import OracleScene
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
    if args.workload == "config5":
        return max(256, args.cpu_sample // 8)
    return args.cpu_sample if args.mode != "live" else max(256, args.cpu_sample // 16)


def scene_name(args):
    return "barrett_mug_200k" if args.workload == "config5" else "barrett_mug"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from graspit_b200.formats.scenepack import load_pack
    from graspit_b200.planner import SearchVariables, SimAnnParams
    from graspit_b200.sampling import sample_planner_states
    from oracle import ref
    scene = load_pack(scene_name(args))
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
        if args.workload == "anneal":
            st = cpu_neighbour_batch(sv, params, sample, step, 1234)
        else:
            st = sample_planner_states(scene, sample, 1234 + step)
        dt = cpu.time(st, mode)
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.workload == "config5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, poses_per_gpu(args, max(1, args.gpus))),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{sample} neighbour states per step of the same workload (reference GraspitCollision compiled "
                                       "verbatim, -O3, one instance per thread; FK precomputed outside the timed call)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def poses_per_gpu(args, world):
    return CONFIG5_TOTAL // world if args.workload == "config5" else args.poses


def workload_config(args, poses, mode=None, workload=None):
    mode = mode or args.mode
    workload = workload or args.workload
    m = "CONTACT_LIVE" if mode == "live" else "CONTACT_PRESET"
    if workload == "anneal":
        w = ("Barrett hand vs mug (worlds/plannerMug.xml), EGPlanner SimAnnPlanner CONTACT_ENERGY " + m +
             f", {poses} annealing chains per GPU, one neighbour per chain per step (SimAnnParams ANNEAL_DEFAULT, "
             "SPACE_AXIS_ANGLE + POSE_EIGEN = 8 fp32 variables, Philox seed 1234, chains started uniformly in the planner's box)")
    elif workload == "config5":
        w = (f"BASELINE.json configs[4]: Barrett hand vs the mug subdivided to 220 800 triangles, {CONFIG5_TOTAL} annealing chains in total "
             f"({poses} per GPU, strong scaling), CONTACT_ENERGY " + m + ", one neighbour per chain per step, chains started on the config-1 "
             "strata (planner box / 150 mm ball), best-row exchange every step")
    else:
        w = ("Barrett hand vs mug (worlds/plannerMug.xml), CONTACT_ENERGY " + m + f", {poses} random planner states per step per GPU "
             "(SPACE_AXIS_ANGLE + POSE_EIGEN, 8 fp32 variables; 50% planner-box / 50% within 150 mm strata)")
    return {"workload": w, "poses_per_step_per_gpu": poses, "mode": mode, "kind": workload,
            "l2": "flushed between timed steps (256 MiB write); per-step inputs 32 B / pose"}


# ------------------------------------------------------------------------------------------------------- product arm
class Bench:
    """torch / distributed plumbing shared by the headline run and the sub-records"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        # one explicit stream for torch ops, the engine's kernels and the timing events
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.synchronize()
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]


def run_workload(B: Bench, scene_id: str, workload: str, npose: int, K: int, W: int, live: bool, exchange: str, e2e_steps: int):
    """One measured workload on this rank's GPU.  Returns a dict of timings (already max-reduced over the ranks) and
    explanatory numbers.  `workload`: anneal | config5 (annealing chains resident on the device) | random (fresh states)."""
    torch, dist = B.torch, B.dist
    from graspit_b200.engine import B2C_EVAL_LIVE, B2C_INPUT_AA_EIGEN, Annealer, Engine, EvalArgs, SceneBinding
    from graspit_b200.formats.scenepack import load_pack
    from graspit_b200.planner import BEST_LIST_SIZE, BestList, PeerExchange, SearchVariables, SimAnnParams
    from graspit_b200.sampling import object_body, sample_planner_states

    rank, world, dev, stream = B.rank, B.world, B.dev, B.stream
    scene = load_pack(scene_id)
    obj = object_body(scene)
    eng = Engine(B.local_rank)
    sb = SceneBinding(eng, scene)
    rid, oid = sb.robot_ids[0], sb.body_ids[obj]
    ref_tran = scene.body_tran[obj]
    neg = scene.robots[0].eg.shape[0]
    nvar = 6 + neg
    sv = SearchVariables.axis_angle_eigen(neg)
    params = SimAnnParams.ANNEAL_DEFAULT()
    eng.set_stream(stream.cuda_stream)
    anneal = workload in ("anneal", "config5")

    # ---- device-resident annealing: chains [rank * npose, (rank + 1) * npose) of world * npose, chains never migrate
    if workload == "config5":
        init = sample_planner_states(scene, npose, 1234 + rank)                    # config-1 strata
    else:
        init = start_states(sv, npose, 1234 + rank)
    an = Annealer(eng, rid, oid, sv, params, init, ref_tran, seed=1234, chain_offset=rank * npose, live=live)
    best_list = BestList(ref_tran)
    # exchange of the best rows: "peer" = stores into the peers' rings by the best-list kernel (no collective);
    # "nccl" = all-gather of the rows (the checked alternative).  At N = 1 the same calls keep the planner's own best list.
    px = PeerExchange(eng, an, rank, world, rows=BEST_LIST_SIZE, depth=4, device=dev)
    RING = 3
    host_rows = [torch.zeros((world, BEST_LIST_SIZE, nvar + 1), dtype=torch.float32).pin_memory() for _ in range(RING)]
    host_ev = [torch.cuda.Event() for _ in range(RING)]
    rows_d = [torch.empty((BEST_LIST_SIZE, nvar + 1), dtype=torch.float32, device=dev) for _ in range(RING)]
    gathered = [torch.empty((world, BEST_LIST_SIZE, nvar + 1), dtype=torch.float32, device=dev) for _ in range(RING)]
    pending = {}
    merged = {"offers": 0}

    def merge_host(slot):
        r = host_rows[slot].numpy().reshape(-1, nvar + 1)
        e = r[:, -1]
        m = e < min(best_list.worst_energy(), 1.0e7)
        if m.any():
            merged["offers"] += int(m.sum())
            best_list.merge(r[m, :-1], e[m])

    def exchange_step(i):
        """issue the exchange of step i; merge (on the host) the rows collected at step i - 1, which have long landed"""
        slot = i % RING
        if exchange == "none":
            return
        if exchange == "peer":
            an.publish()
            eng.exchange_collect(host_rows[slot].data_ptr(), 0, True)
        else:
            an.best_rows(BEST_LIST_SIZE, rows_d[slot].data_ptr())
            if world > 1:
                # asynchronous all-gather of this step's rows; the rows gathered at the PREVIOUS step go to the host now
                # (the stream waits for that gather only: one step of slack between the ranks)
                pending[i] = dist.all_gather_into_tensor(gathered[slot].view(-1, nvar + 1), rows_d[slot], async_op=True)
                if (i - 1) in pending:
                    pending.pop(i - 1).wait()
                    host_rows[slot].copy_(gathered[(i - 1) % RING], non_blocking=True)
            else:
                host_rows[slot][0].copy_(rows_d[slot], non_blocking=True)
        host_ev[slot].record(stream)
        prev = (i - 1) % RING
        if i >= 1:
            host_ev[prev].synchronize()                    # one step behind the GPU: never idles it
            merge_host(prev)

    # host copies of the neighbour batches of the end-to-end leg, taken from the same chains
    host_batches = []
    if anneal:
        for _ in range(e2e_steps + min(W, 3)):
            host_batches.append(an.propose())
            an.step(1)
    else:
        host_batches = [sample_planner_states(scene, npose, 1234 + 1000 * rank + s) for s in range(K + W)]
    dev_batches = None if anneal else [torch.from_numpy(b).to(dev) for b in host_batches]
    legal_d = torch.zeros(npose, dtype=torch.uint8, device=dev)
    energy_d = torch.zeros(npose, dtype=torch.float32, device=dev)
    eng.set_profiling(True)

    def step_device(i):
        if anneal:
            an.step(1)
            exchange_step(i)
        else:
            st = dev_batches[i]
            eng.eval_batch_device(rid, oid, st.data_ptr(), npose, nvar, legal_d.data_ptr(), energy_d.data_ptr(),
                                  input_mode=B2C_INPUT_AA_EIGEN, ref_tran=ref_tran, live=live)

    def drain():
        for h in list(pending.values()):
            h.wait()
        pending.clear()
        B.barrier()

    for i in range(W):
        step_device(i)
    drain()

    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kernel_ms = []
    drain()
    for s in range(K):
        B.flush.fill_(s & 0xff)                     # L2 flush, outside the event bracket
        ev[s][0].record(stream)
        step_device(W + s)
        ev[s][1].record(stream)
        kernel_ms.append(eng.last_kernel_ms())      # synchronises; after the bracket
    drain()
    eng.check_overflow()
    launches = eng.launch_count - launches0
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    stats = eng.last_stats()
    d = an.read() if anneal else None
    anneal_info = None
    if anneal:
        anneal_info = {"steps_taken": d["steps"], "neighbours_evaluated": d["evaluated"], "legal_fraction": d["legal_neighbours"] / max(1, d["evaluated"]),
                       "jump_fraction": d["jumps"] / max(1, d["evaluated"]), "best_energy_min": float(d["best_energy"].min()),
                       "best_energy_median": float(np.median(d["best_energy"])), "best_list": len(best_list.states),
                       "best_list_energy": [round(float(x), 4) for x in best_list.energies[:3]], "rows_offered_to_best_list": merged["offers"]}

    # ---- end to end through the C ABI with host buffers: H2D + kernels + D2H + exchange inside the timed region
    eng.set_profiling(False)
    KE = min(e2e_steps, len(host_batches) - min(W, 3)) if anneal else min(e2e_steps, K)
    WE = min(W, 3)

    def e2e_leg(pinned: bool):
        if pinned:
            bufs = [torch.from_numpy(b).pin_memory() for b in host_batches[: WE + KE]]
            lg = torch.zeros(npose, dtype=torch.uint8).pin_memory()
            en = torch.zeros(npose, dtype=torch.float32).pin_memory()
        else:                                       # what a GraspIt!-side std::vector is: pageable
            bufs = [torch.from_numpy(np.array(b, copy=True)) for b in host_batches[: WE + KE]]
            lg = torch.zeros(npose, dtype=torch.uint8)
            en = torch.zeros(npose, dtype=torch.float32)

        def step_host(i):
            a = EvalArgs()
            a.robot, a.object, a.npose, a.input_mode = rid, oid, npose, B2C_INPUT_AA_EIGEN
            a.states, a.stride = bufs[i].data_ptr(), nvar
            a.ref_q[:] = list(ref_tran[0]); a.ref_t[:] = list(ref_tran[1])
            a.flags = B2C_EVAL_LIVE if live else 0
            a.legal, a.energy = lg.data_ptr(), en.data_ptr()
            eng._ck(eng.L.b2c_eval_batch(eng.h, C.byref(a)))
            if anneal:
                # the planner's exchange belongs to the step: publish + collect + merge, synchronously
                an.publish()
                eng.exchange_collect(host_rows[0].data_ptr(), 0, False)
                merge_host(0)

        for i in range(WE):
            step_host(i)
        B.barrier()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KE)]
        for s in range(KE):
            B.flush.fill_(s & 0xff)
            ev2[s][0].record(stream)
            step_host(WE + s)
            ev2[s][1].record(stream)
        B.barrier()
        return float(sum(a.elapsed_time(b) for a, b in ev2)), int(lg.sum().item())

    e2e_ms, nlegal = e2e_leg(True)
    e2e_pg_ms, _ = e2e_leg(False)
    total_ms, e2e_ms, e2e_pg_ms = B.max_over_ranks([total_ms, e2e_ms, e2e_pg_ms])
    kms = {k: float(np.mean([x[k] for x in kernel_ms])) for k in kernel_ms[0]}
    out = dict(ms_per_step=total_ms / K, value=world * npose / (total_ms / K / 1e3), kernel_ms=kms, launches=int(launches), stats=stats,
               anneal=anneal_info, e2e_ms=e2e_ms / KE, e2e_value=world * npose / (e2e_ms / KE / 1e3), e2e_steps=KE,
               e2e_pageable_value=world * npose / (e2e_pg_ms / KE / 1e3), nlegal_last=nlegal, npose=npose, nvar=nvar, steps=K,
               scene=scene, host_batches=host_batches, WE=WE, triangles=int(sum(len(b.tris) for b in scene.bodies)))
    an.close()
    eng.exchange_destroy()
    eng.close()
    return out


def roofline(res, mode, root):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kms, npose, nvar, nlegal = res["kernel_ms"], res["npose"], res["nvar"], res["nlegal_last"]
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
        tr = json.load(open(os.path.join(root, "profiles", "traffic.json")))
        traffic = tr.get(mode, {}).get(f"{dom}_kernel")
    except Exception:
        pass
    return {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": algo_bytes,
            "kernel_ms_all": kms, "legal_fraction_last_step": nlegal / npose,
            "note": "working set is L2-resident by design: compulsory HBM traffic only (SURVEY.md 8(d)); the kernel is "
                    "bound by instruction issue under divergence, see profiles/; traffic = dram bytes of one ncu "
                    "capture (profiles/traffic.json); traversal work of the last step: " + json.dumps(res["stats"])}


# ------------------------------------------------------------------------------------- secondary configurations (N = 1)
def _rest_postures(eng, sb, sc, n, seed, strata=("near",)):
    """n collision-free hand poses near the object with the DOFs at rest: RAW rows (base q, t + DOFs).  The base poses come
    from the engine's own state mapping (body_tf of an AA + EIGEN evaluation), so nothing outside the product is involved."""
    from graspit_b200.engine import B2C_INPUT_AA_EIGEN
    from graspit_b200.sampling import hand_robot, object_body, sample_aa_states
    ridx = hand_robot(sc)
    r = sc.robots[ridx]
    rid, obj = sb.robot_ids[ridx], sb.body_ids[object_body(sc)]
    rest = np.clip(r.dof_default, r.dof_min, r.dof_max).astype(np.float32)
    nlink = sum(len(c.link_bodies) for c in r.chains)
    raw = np.zeros((0, 7 + r.n_dof), dtype=np.float32)
    while len(raw) < n:
        pos, _ = sample_aa_states(sc, 4 * n, seed=seed, strata=strata)
        st = np.concatenate([pos, np.zeros((len(pos), r.eg.shape[0]), np.float32)], 1).astype(np.float32)
        tf = eng.eval_batch(rid, obj, st, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[object_body(sc)], body_tf=True)["body_tf"]
        cand = np.concatenate([tf[:, nlink, :].astype(np.float32), np.tile(rest[None, :], (len(pos), 1))], 1)
        raw = np.concatenate([raw, cand[eng.eval_batch(rid, obj, cand)["legal"] == 1]])[:n]
        seed += 1
    return raw


def secondary_configs(device):
    """BASELINE.json configs[2] (contacts of closed hands), configs[3] (arm + hand collisions) and the batched autograsp as
    throughput sub-records: one timed call each after a warm-up call of the same size, host buffers, rank 0 at N = 1 only."""
    import time
    from graspit_b200.engine import AG_INTERP_ERROR, B2C_INPUT_JOINTS, Engine, SceneBinding
    from graspit_b200.formats.scenepack import load_pack
    out = {}
    # ---- autograsp: Hand::autoGrasp to contact for 65 536 Barrett poses in one call
    sc = load_pack("barrett_mug")
    eng = Engine(device)
    sb = SceneBinding(eng, sc)
    rid = sb.robot_ids[0]
    raw = _rest_postures(eng, sb, sc, 65536, 7)
    raw[:, 8:] = 0.0                                       # fingers open, spread at rest
    eng.auto_grasp(rid, sc.robots[0], raw)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    ag = eng.auto_grasp(rid, sc.robots[0], raw)
    dt = time.perf_counter() - t0
    out["autograsp"] = {"metric": "autograsps/s (Hand::autoGrasp to contact, Barrett vs mug, one b2c_autograsp_batch call, host buffers)",
                        "npose": len(raw), "seconds": dt, "value": len(raw) / dt, "unit": "autograsps/s", "rounds": int(ag["rounds"]),
                        "gpu_launches": int(eng.launch_count - l0), "mean_steps": float(ag["iters"].mean()),
                        "errors": int(((ag["status"] & AG_INTERP_ERROR) != 0).sum())}
    eng.close()
    # ---- config 3: HumanHand20DOF vs flask, hands closed by the autograsp, allContacts(0.2) per pose in one call
    sc = load_pack("humanhand_flask")
    eng = Engine(device)
    sb = SceneBinding(eng, sc)
    rid = sb.robot_ids[0]
    raw = _rest_postures(eng, sb, sc, 16384, 100)
    ag = eng.auto_grasp(rid, sc.robots[0], raw)
    ok = (ag["status"] & AG_INTERP_ERROR) == 0
    js = np.concatenate([raw[ok, :7].astype(np.float64), ag["joints"][ok]], axis=1)
    eng.contacts_batch(rid, js, 0.2, input_mode=B2C_INPUT_JOINTS)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    res = eng.contacts_batch(rid, js, 0.2, input_mode=B2C_INPUT_JOINTS)
    dt = time.perf_counter() - t0
    out["config3"] = {"metric": "poses/s: allContacts(0.2) per pose, HumanHand20DOF vs flask (BASELINE.json configs[2]), one b2c_contacts_batch call, host buffers",
                      "npose": len(js), "seconds": dt, "value": len(js) / dt, "unit": "poses/s", "contacts_total": int(len(res["contacts"])),
                      "contacts_per_pose": float(len(res["contacts"]) / max(1, len(js))), "active_pairs": len(eng.robot_pairs(rid)),
                      "gpu_launches": int(eng.launch_count - l0)}
    # ... and the neighbourhoods of both sides of every contact on an elastic body (findSoftNeighborhoods -> bodyRegion), one call
    from graspit_b200 import engine as E
    pairs = eng.robot_pairs(rid)
    inv_b = {v: k for k, v in sb.body_ids.items()}
    y = np.array([[sc.bodies[inv_b[a]].youngs, sc.bodies[inv_b[b]].youngs] for a, b in pairs])
    soft = y.max(axis=1)[res["pair"]] > 0
    if soft.any():
        cs, pq = res["contacts"][soft], res["pair"][soft]
        pa = np.array(pairs)[pq]
        rad = np.array([E.soft_neighborhood_radius(*y[q]) for q in pq])
        qb = np.stack([pa[:, 0], pa[:, 1]], 1).reshape(-1)
        qp = np.stack([cs[:, 0:3], cs[:, 3:6]], 1).reshape(-1, 3)
        qn = np.stack([cs[:, 6:9], cs[:, 9:12]], 1).reshape(-1, 3)
        qr = np.repeat(rad, 2)
        eng.body_region_batch(qb, qp, qn, qr, flat=True)
        l0 = eng.launch_count
        t0 = time.perf_counter()
        flat, off = eng.body_region_batch(qb, qp, qn, qr, flat=True)
        dt = time.perf_counter() - t0
        out["config3"]["region_batch"] = {"metric": "bodyRegion queries/s (both sides of every contact on an elastic body, one b2c_region_batch call, host buffers)",
                                          "queries": int(len(qb)), "seconds": dt, "value": len(qb) / dt, "unit": "queries/s",
                                          "points_out": int(len(flat)), "gpu_launches": int(eng.launch_count - l0)}
    eng.close()
    # ---- config 4: Staubli TX60L + Barrett vs obstacles, allCollisions(ALL) for 262 144 arm + hand configurations
    sc = load_pack("staubli_barrett")
    eng = Engine(device)
    sb = SceneBinding(eng, sc)
    arm, hand = sc.robots[0], sc.robots[1]
    rng = np.random.default_rng(1234)
    n = 262144
    base = np.tile(np.concatenate([arm.tran[0], arm.tran[1]])[None, :], (n, 1))
    dofs = np.concatenate([rng.uniform(arm.dof_min, arm.dof_max, size=(n, arm.n_dof)), rng.uniform(hand.dof_min, hand.dof_max, size=(n, hand.n_dof))], 1)
    raw = np.concatenate([base, dofs], 1).astype(np.float32)
    rid = sb.robot_ids[0]
    eng.eval_batch(rid, -1, raw, pair_mask=True)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    res = eng.eval_batch(rid, -1, raw, pair_mask=True)
    dt = time.perf_counter() - t0
    out["config4"] = {"metric": "configurations/s: allCollisions(ALL, interest = robot), Staubli TX60L + Barrett vs obstacles (BASELINE.json configs[3]), host buffers",
                      "n": n, "seconds": dt, "value": n / dt, "unit": "configurations/s", "collision_free": float(res["legal"].mean()),
                      "active_pairs": len(eng.robot_pairs(rid)), "triangles": int(sum(len(b.tris) for b in sc.bodies)),
                      "gpu_launches": int(eng.launch_count - l0)}
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="preset", choices=["preset", "live"])
    ap.add_argument("--workload", default="anneal", choices=["anneal", "random", "config5"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl", "none"], help="none: no best-list upkeep at all (diagnostic)")
    ap.add_argument("--poses", type=int, default=65536)
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the LIVE and config-5 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    B = Bench(args)
    world, rank = B.world, B.rank
    live = args.mode == "live"
    npose = poses_per_gpu(args, world)
    K, W = args.steps, args.warmup
    if args.workload == "config5":
        K, W = min(K, 10), min(W, 3)                 # 1 M chains: a step is ~0.1 s of kernels at N = 1

    sampler = ClockSampler(B.local_rank)
    sampler.start()
    res = run_workload(B, scene_name(args), args.workload, npose, K, W, live, args.exchange, e2e_steps=min(K, 20))
    clocks = sampler.stop()

    sub = {}
    if not args.no_sub and args.workload == "anneal" and not live:
        r = run_workload(B, "barrett_mug", "anneal", npose, 3, 3, True, args.exchange, e2e_steps=3)
        sub["live"] = {"config": workload_config(args, npose, mode="live"), "steps": 3, "warmup": 3, "ms_per_step": r["ms_per_step"], "value": r["value"],
                       "unit": UNIT, "e2e": {"value": r["e2e_value"], "pageable_value": r["e2e_pageable_value"]}, "kernel_ms": r["kernel_ms"],
                       "gpu_launches": r["launches"], "traversal": r["stats"]}
        n5 = CONFIG5_TOTAL // world
        r = run_workload(B, "barrett_mug_200k", "config5", n5, 3, 3, False, args.exchange, e2e_steps=3)
        sub["config5"] = {"config": workload_config(args, n5, mode="preset", workload="config5"), "steps": 3, "warmup": 3, "scaling": "strong",
                          "ms_per_step": r["ms_per_step"], "value": r["value"], "unit": UNIT,
                          "e2e": {"value": r["e2e_value"], "pageable_value": r["e2e_pageable_value"]}, "kernel_ms": r["kernel_ms"],
                          "gpu_launches": r["launches"], "object_triangles": 220800, "legal_fraction": r["anneal"]["legal_fraction"], "traversal": r["stats"]}

    if not args.no_sub and world == 1 and args.workload == "anneal" and not live:
        try:
            sub.update(secondary_configs(B.local_rank))
        except Exception as ex:          # secondary numbers never break the headline line
            sub["secondary_error"] = str(ex)[:200]

    if rank == 0:
        nvar = res["nvar"]
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": res["steps"], "warmup": W,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if args.workload == "config5" else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, npose),
                "e2e": {"value": res["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": npose * nvar * 4, "d2h_bytes_per_step": npose * 5 + world * 20 * (nvar + 1) * 4,
                        "ms_per_step": res["e2e_ms"], "steps": res["e2e_steps"], "host_buffers": "pinned", "pageable_value": res["e2e_pageable_value"],
                        "includes": "b2c_eval_batch (H2D states, kernels, D2H legal + energy) + best-row exchange (publish, collect, host merge)"},
                "gpu_launches": res["launches"], "clocks": clocks, "exchange": args.exchange,
                "roofline": roofline(res, args.mode, ROOT)}
        if res["anneal"]:
            line["anneal"] = res["anneal"]
        line.update(sub)
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref
                threads = max(1, ref.hardware_threads())
                sample = cpu_sample_size(args)
                cpu = CpuReference(res["scene"], threads)
                hb = res["host_batches"][res["WE"]]
                cpu.time(hb[: max(64, sample // 8)], 1 if live else 0)            # warm-up
                secs = min(cpu.time(hb[:sample], 1 if live else 0) for _ in range(3))
                line["cpu_baseline"] = {"value": sample / secs, "unit": UNIT, "cores": threads, "kind": "reference",
                                        "sample": f"first {sample} states of the first timed end-to-end batch, best of 3 ({secs:.2f} s); reference "
                                                  "GraspitCollision compiled verbatim (-O3), one instance per thread, FK precomputed"}
            except Exception as ex:   # the oracle is a checker, never a dependency of the measurement
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}
        print(json.dumps(line))
    if world > 1:
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
