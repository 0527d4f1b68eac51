"""Host side of the batched EGPlanner / SimAnnPlanner: search-variable tables, SimAnn parameter presets, the
best-list merge and the multi-GPU exchange of per-rank best states.

Mirrors, with the reference's names and values:
  * ``PositionStateAA::createVariables`` / ``PostureStateEigen::createVariables``
    (reference src/EGPlanner/searchStateImpl.cpp:50-80,147-155)
  * ``SimAnnParams`` presets (src/EGPlanner/simAnnParams.cpp)
  * ``SimAnnPlanner::mainLoop`` best-list rule and ``EGPlanner::addToListOfUniqueSolutions``
    (src/EGPlanner/simAnnPlanner.cpp:111-148, egPlanner.cpp:527-557), ``HandObjectState::distance``
    (src/EGPlanner/searchState.cpp:572-611)
The annealing itself runs on the device (``engine.Annealer``); nothing here touches geometry.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

BEST_LIST_SIZE = 20          # simAnnPlanner.cpp:36


@dataclass
class SimAnnParams:
    YC: float; HC: float; YDIMS: int; HDIMS: int; NBR_ADJ: float; ERR_ADJ: float; T0: float; K0: int

    @staticmethod
    def ANNEAL_DEFAULT():
        return SimAnnParams(7.0, 7.0, 8, 8, 1.0, 1.0e-6, 1.0e6, 30000)

    @staticmethod
    def ANNEAL_LOOP():
        return SimAnnParams(7.0, 7.0, 8, 8, 1.0, 1.0e-6, 1.0e6, 35000)

    @staticmethod
    def ANNEAL_MODIFIED():
        return SimAnnParams(2.38, 2.38, 8, 8, 1.0e-3, 1.0e-1, 1.0e3, 0)

    @staticmethod
    def ANNEAL_STRICT():
        return SimAnnParams(0.72, 0.22, 2, 2, 1.0, 1.0, 10.0, 0)

    @staticmethod
    def ANNEAL_ONLINE():
        return SimAnnParams(0.24, 0.54, 2, 3, 1.0, 1.0e-2, 1.0e1, 100)


@dataclass
class SearchVariables:
    names: List[str]
    vmin: np.ndarray
    vmax: np.ndarray
    default: np.ndarray
    vjump: np.ndarray
    circular: np.ndarray

    @staticmethod
    def axis_angle_eigen(n_eg: int, eg_limits=None) -> "SearchVariables":
        """SPACE_AXIS_ANGLE position + POSE_EIGEN posture: the planner dialog's default state layout."""
        pi = np.pi
        rows = [("Tx", -250, 250, 0, 150, 0), ("Ty", -250, 250, 0, 150, 0), ("Tz", -250, 350, 350, 150, 0),
                ("theta", 0, pi, 0, pi / 5, 0), ("phi", -pi, pi, 0, pi / 2, 1), ("alpha", 0, pi, pi / 2, pi / 2, 0)]
        for i in range(n_eg):
            lo, hi = -4.0, 4.0
            if eg_limits is not None and not np.isnan(eg_limits[i][0]):
                lo, hi = float(eg_limits[i][0]), float(eg_limits[i][1])
            rows.append((f"EG {i}", lo, hi, lo, (hi - lo) / 4.0, 0))
        a = np.array([r[1:] for r in rows], dtype=np.float64)
        return SearchVariables([r[0] for r in rows], a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy(), a[:, 3].copy(),
                               a[:, 4].astype(np.uint8))


def _quat_mul(a, b):
    w1, x1, y1, z1 = a; w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def state_distance(pose_a, pose_b, max_radius: float) -> float:
    """HandObjectState::distance (searchState.cpp:572-611) on hand poses already multiplied by the approach
    transform: max(|dt| / maxRadius, 0.5 |angle| / pi).  pose = (q wxyz, t)."""
    d = np.linalg.norm(np.asarray(pose_a[1]) - np.asarray(pose_b[1])) / max_radius
    qa, qb = np.asarray(pose_a[0], dtype=np.float64), np.asarray(pose_b[0], dtype=np.float64)
    q = _quat_mul(qa, np.array([qb[0], -qb[1], -qb[2], -qb[3]]))
    q = q / np.linalg.norm(q)
    if q[0] < 0:                      # Eigen::AngleAxisd(Quaternion): angle = 2 acos(w) after flipping w < 0
        q = -q
    angle = 2.0 * np.arctan2(np.linalg.norm(q[1:]), q[0])
    return max(d, 0.5 * abs(angle) / np.pi)


def aa_state_pose(state, ref_tran):
    """total hand transform % approach, i.e. mRefTran % T(tx,ty,tz) % AxisAngle(alpha, axis(theta, phi)): the
    approach factors cancel (searchStateImpl.cpp:156-168, searchState.cpp:592)."""
    tx, ty, tz, th, ph, al = [float(x) for x in state[:6]]
    axis = np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
    if al * al < 1.0e-6:              # transf::AXIS_ANGLE_ROTATION snaps tiny angles to identity (transform.cpp:26-29)
        q = np.array([1.0, 0, 0, 0])
    else:
        q = np.concatenate([[np.cos(al / 2)], np.sin(al / 2) * axis / np.linalg.norm(axis)])
    rq, rt = np.asarray(ref_tran[0], dtype=np.float64), np.asarray(ref_tran[1], dtype=np.float64)
    w, v = rq[0], rq[1:]
    t = np.array([tx, ty, tz])
    rot_t = t + 2.0 * w * np.cross(v, t) + 2.0 * np.cross(v, np.cross(v, t))
    return _quat_mul(rq, q), rt + rot_t


def _poses_of(states, ref_tran):
    """aa_state_pose for a batch: (q (n,4), t (n,3))"""
    st = np.asarray(states, dtype=np.float64).reshape(-1, np.asarray(states).shape[-1])
    out_q = np.zeros((len(st), 4)); out_t = np.zeros((len(st), 3))
    for i, s in enumerate(st):
        out_q[i], out_t[i] = aa_state_pose(s, ref_tran)
    return out_q, out_t


class BestList:
    """mBestList of SimAnnPlanner: at most BEST_LIST_SIZE states, sorted by energy, no two closer than 0.2.
    The poses of the listed states are cached, and a candidate is compared with the whole list in one numpy expression
    (HandObjectState::distance, searchState.cpp:572-611): the merge runs inside the timed planner step."""

    def __init__(self, ref_tran, max_radius=200.0, size=BEST_LIST_SIZE, distance=0.2):
        self.ref_tran, self.max_radius, self.size, self.distance = ref_tran, max_radius, size, distance
        self.states: List[np.ndarray] = []
        self.energies: List[float] = []
        self._q = np.zeros((0, 4)); self._t = np.zeros((0, 3))
        self._seen = set()

    def worst_energy(self) -> float:
        return 1.0e5 if len(self.states) < self.size else self.energies[-1]      # simAnnPlanner.cpp:128-130

    def _distances(self, pose):
        """state_distance(pose, every listed pose)"""
        if not self.states:
            return np.zeros(0)
        d = np.linalg.norm(self._t - pose[1][None, :], axis=1) / self.max_radius
        qa = np.asarray(pose[0], dtype=np.float64)
        qb = self._q
        # q = qa * conj(qb), normalised, w >= 0; angle = 2 atan2(|xyz|, w)
        w = qa[0] * qb[:, 0] + qa[1] * qb[:, 1] + qa[2] * qb[:, 2] + qa[3] * qb[:, 3]
        x = -qa[0] * qb[:, 1] + qa[1] * qb[:, 0] - qa[2] * qb[:, 3] + qa[3] * qb[:, 2]
        y = -qa[0] * qb[:, 2] + qa[1] * qb[:, 3] + qa[2] * qb[:, 0] - qa[3] * qb[:, 1]
        z = -qa[0] * qb[:, 3] - qa[1] * qb[:, 2] + qa[2] * qb[:, 1] + qa[3] * qb[:, 0]
        n = np.sqrt(w * w + x * x + y * y + z * z)
        w, x, y, z = w / n, x / n, y / n, z / n
        sgn = np.where(w < 0, -1.0, 1.0)
        angle = 2.0 * np.arctan2(np.sqrt(x * x + y * y + z * z), w * sgn)
        return np.maximum(d, 0.5 * np.abs(angle) / np.pi)

    def offer(self, state, energy) -> bool:
        """the JUMP branch of SimAnnPlanner::mainLoop (simAnnPlanner.cpp:131-143)"""
        if not energy < self.worst_energy():
            return False
        pose = aa_state_pose(state, self.ref_tran)
        # addToListOfUniqueSolutions (egPlanner.cpp:527-557): walk the list in order; a near entry with higher energy is
        # erased, the first near entry with lower-or-equal energy rejects the candidate (entries erased before it stay erased)
        near = np.abs(self._distances(pose)) < self.distance
        keep = np.ones(len(self.states), dtype=bool)
        add = True
        for i in np.nonzero(near)[0]:
            if energy < self.energies[i]:
                keep[i] = False
            else:
                add = False
                break
        if not keep.all():
            self.states = [s for s, k in zip(self.states, keep) if k]
            self.energies = [e for e, k in zip(self.energies, keep) if k]
            self._q, self._t = self._q[keep], self._t[keep]
        if not add:
            return False
        self.states.append(np.array(state, dtype=np.float32)); self.energies.append(float(energy))
        self._q = np.concatenate([self._q, np.asarray(pose[0], dtype=np.float64)[None]], 0)
        self._t = np.concatenate([self._t, np.asarray(pose[1], dtype=np.float64)[None]], 0)
        order = np.argsort(np.array(self.energies), kind="stable")[: self.size]
        self.states = [self.states[j] for j in order]
        self.energies = [self.energies[j] for j in order]
        self._q, self._t = self._q[order], self._t[order]
        return True

    def merge(self, states: Sequence, energies: Sequence):
        """offer candidates in ascending energy (the order a single planner would have met the better ones).  Rows that were
        offered before (the exchange repeats a GPU's best rows until they change) are skipped: offering a state twice
        cannot change the list."""
        e = np.asarray(energies, dtype=np.float64)
        st = np.asarray(states, dtype=np.float32)
        for j in np.argsort(e, kind="stable"):
            if not e[j] < 1.0e7:
                continue
            key = (st[j].tobytes(), float(e[j]))
            if key in self._seen:
                continue
            if len(self._seen) > 100000:
                self._seen.clear()
            self._seen.add(key)
            self.offer(st[j], float(e[j]))


def complete_dof_to_raw(states, ref_tran):
    """SPACE_COMPLETE + POSE_DOF rows (Tx Ty Tz Qw Qx Qy Qz, DOFs) -> RAW engine rows (q wxyz, t, DOFs) with
    base = mRefTran % transf(Q, T) (HandObjectState::execute, searchState.cpp:515-527; getCoreTran, searchStateImpl.cpp:125-135)"""
    st = np.asarray(states, dtype=np.float64)
    rq, rt = np.asarray(ref_tran[0], dtype=np.float64), np.asarray(ref_tran[1], dtype=np.float64)
    out = np.zeros((st.shape[0], st.shape[1]), dtype=np.float64)
    for i, r in enumerate(st):
        q = r[3:7] / np.linalg.norm(r[3:7])
        w, v = rq[0], rq[1:]
        t = r[0:3]
        rot_t = t + 2.0 * w * np.cross(v, t) + 2.0 * np.cross(v, np.cross(v, t))
        qq = _quat_mul(rq, q)
        out[i, 0:4] = qq / np.linalg.norm(qq); out[i, 4:7] = rt + rot_t; out[i, 7:] = r[7:]
    return out.astype(np.float32)


class ListPlanner:
    """Batched ListPlanner (reference src/EGPlanner/listPlanner.cpp:126-170): every input state is evaluated (one
    b2c_eval_batch call instead of one analyzeState per planner step) and the legal ones enter the best list in input
    order: kept if the list has room or they beat its worst entry; sorted by energy; at most BEST_LIST_SIZE.  Unlike
    SimAnnPlanner no similarity pruning is applied (the reference's ListPlanner does none)."""

    def __init__(self, engine, robot_id: int, object_id: int, ref_tran, live=False, size=BEST_LIST_SIZE):
        self.engine, self.robot, self.object, self.ref_tran, self.live, self.size = engine, robot_id, object_id, ref_tran, live, size
        self.best = []           # (energy, input index)
        self.legal = self.energy = None

    def run(self, states):
        from .engine import B2C_INPUT_AA_EIGEN, B2C_INPUT_RAW
        from .formats.statefile import to_engine_batch
        kind, arr = to_engine_batch(states)
        if len(arr) == 0:
            self.legal, self.energy, self.best = np.zeros(0, np.uint8), np.zeros(0, np.float32), []
            return self.best
        if kind == "aa_eigen":
            r = self.engine.eval_batch(self.robot, self.object, arr, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=self.ref_tran, live=self.live)
        else:
            r = self.engine.eval_batch(self.robot, self.object, complete_dof_to_raw(arr, self.ref_tran), input_mode=B2C_INPUT_RAW, live=self.live)
        self.legal, self.energy = r["legal"], r["energy"]
        best = []
        for i in np.nonzero(self.legal)[0]:
            worst = 1.0e5 if len(best) < self.size else best[-1][0]
            if self.energy[i] < worst:
                best.append((float(self.energy[i]), int(i)))
                best.sort(key=lambda x: x[0])          # std::list::sort is stable
                best = best[: self.size]
        self.best = best
        return best


def local_best_rows(best_states, best_energy, k=BEST_LIST_SIZE):
    """this rank's k best (state || energy) rows as one torch tensor [k, nvar + 1]; works on CPU and CUDA tensors"""
    import torch
    k = min(k, best_energy.numel())
    val, idx = torch.topk(best_energy, k, largest=False)
    return torch.cat([best_states[idx], val[:, None]], 1).contiguous()


def exchange_best(rows, group=None):
    """all-gather of every rank's best rows: the ONLY collective of the path.  NCCL for CUDA tensors,
    gloo for CPU tensors; returns [world * k, nvar + 1]."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    out = torch.empty((world * rows.shape[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, rows, group=group)
    return out


class PeerExchange:
    """The planner's multi-GPU exchange without a collective (b2c_exchange_*): every rank stores its BEST_LIST_SIZE best
    (state || energy) rows straight into its peers' HBM rings over NVLink right after a step (one fused kernel, no NCCL
    launch, no rendezvous); ``collect`` reads whatever has landed.  torch.distributed is used ONCE, to hand the 64-byte
    CUDA IPC handles round (one rank per process); ``peers`` = the other engines when all contexts live in this process."""

    def __init__(self, engine, annealer, rank: int, world: int, rows: int = BEST_LIST_SIZE, depth: int = 4, peers=None, device=None):
        import torch
        self.engine, self.annealer, self.rank, self.world, self.rows = engine, annealer, rank, world, rows
        self.row_floats = annealer.nvar + 1
        handle, self.ring = engine.exchange_create(rank, world, rows, self.row_floats, depth)
        self._handle = handle
        if peers is None and world > 1:
            import torch.distributed as dist
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            allh = torch.empty(world * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, mine)
            engine.exchange_connect(ipc_handles=bytes(allh.cpu().numpy().tobytes()))
        self.host_rows = torch.zeros((world, rows, self.row_floats), dtype=torch.float32).pin_memory() if torch.cuda.is_available() else torch.zeros((world, rows, self.row_floats))
        self.host_steps = torch.zeros(world, dtype=torch.int32).pin_memory() if torch.cuda.is_available() else torch.zeros(world, dtype=torch.int32)

    @staticmethod
    def connect_in_process(exchanges):
        """all contexts in one process (one per device): plain peer access instead of IPC"""
        rings = [x.ring for x in exchanges]
        devs = [x.engine.device for x in exchanges]
        for x in exchanges:
            x.engine.exchange_connect(peer_rings=rings, peer_devices=devs)

    def publish(self):
        self.annealer.publish()

    def collect(self, asynchronous=False):
        """rows [world, rows, nvar + 1] and steps [world] (host tensors; with asynchronous=True valid after a stream sync)"""
        self.engine.exchange_collect(self.host_rows.data_ptr(), self.host_steps.data_ptr(), asynchronous)
        return self.host_rows, self.host_steps

    def merge_into(self, best_list):
        """offer the collected rows to a BestList (only rows that beat its current worst entry reach the distance test)"""
        r = self.host_rows.numpy().reshape(-1, self.row_floats)
        e = r[:, -1]
        m = e < min(best_list.worst_energy(), 1.0e7)
        if m.any():
            best_list.merge(r[m, :-1], e[m])


def shard(nchain_total: int, rank: int, world: int):
    """block partition of the chains over ranks: chains [lo, hi) live on `rank` and never migrate"""
    per = nchain_total // world
    rem = nchain_total % world
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)
