"""GraspIt! planner-state text files: the wire format of ``HandObjectState::writeToFile / readFromFile``
(reference src/EGPlanner/searchState.cpp:480-513) built from ``VariableSet::writeToFile / readFromFile``
(:224-262).  One state = two lines, posture first:

    <posture type> v0 v1 ...      POSE_EIGEN = 4 (eigengrasp amplitudes) or POSE_DOF = 5 (raw DOF values)
    <position type> v0 v1 ...     SPACE_COMPLETE = 0 (Tx Ty Tz Qw Qx Qy Qz), SPACE_AXIS_ANGLE = 1 (Tx Ty Tz theta phi alpha),
                                  SPACE_ELLIPSOID = 2 (beta gamma tau dist), SPACE_APPROACH = 3 (dist wrist1 wrist2)

Values are printed with C ``%f`` (six decimals) followed by a space, exactly as the reference does; it is the input
format of ``ListPlanner`` (src/EGPlanner/listPlanner.cpp) and the natural interchange format for pose batches.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List

import numpy as np

SPACE_COMPLETE, SPACE_AXIS_ANGLE, SPACE_ELLIPSOID, SPACE_APPROACH, POSE_EIGEN, POSE_DOF = range(6)   # search.h:6-8
_POSITION_SIZES = {SPACE_COMPLETE: 7, SPACE_AXIS_ANGLE: 6, SPACE_ELLIPSOID: 4, SPACE_APPROACH: 3}


@dataclass
class PlannerState:
    posture_type: int
    posture: np.ndarray
    position_type: int
    position: np.ndarray


def _line(kind: int, values) -> str:
    return "%d " % kind + "".join("%f " % float(v) for v in values) + "\n"


def write_states(path_or_file, states: Iterable[PlannerState]):
    f = open(path_or_file, "w") if isinstance(path_or_file, str) else path_or_file
    try:
        for s in states:
            f.write(_line(s.posture_type, s.posture))
            f.write(_line(s.position_type, s.position))
    finally:
        if isinstance(path_or_file, str):
            f.close()


def read_states(path_or_file, n_posture: int | None = None) -> List[PlannerState]:
    """n_posture: number of posture variables expected (the hand's eigengrasp / DOF count); None = whatever the line holds.
    A malformed state ends the read, as the reference's loader stops at the first state it cannot parse."""
    f = open(path_or_file) if isinstance(path_or_file, str) else path_or_file
    try:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    finally:
        if isinstance(path_or_file, str):
            f.close()
    out = []
    for i in range(0, len(lines) - 1, 2):
        a, b = lines[i].split(), lines[i + 1].split()
        try:
            pt, st = int(a[0]), int(b[0])
            pv = np.array([float(x) for x in a[1:]], dtype=np.float64)
            sv = np.array([float(x) for x in b[1:]], dtype=np.float64)
        except (ValueError, IndexError):
            break
        if pt not in (POSE_EIGEN, POSE_DOF) or st not in _POSITION_SIZES:
            break
        if len(sv) < _POSITION_SIZES[st] or (n_posture is not None and len(pv) < n_posture):
            break                                        # "Line to short to read all state variables"
        sv = sv[:_POSITION_SIZES[st]]
        if n_posture is not None:
            pv = pv[:n_posture]
        out.append(PlannerState(pt, pv, st, sv))
    return out


def to_engine_batch(states: List[PlannerState]):
    """states of one layout -> (input_mode name, float32 array) for Engine.eval_batch.
    SPACE_AXIS_ANGLE + POSE_EIGEN -> ("aa_eigen", [n, 6 + nEG]); SPACE_COMPLETE + POSE_DOF -> ("complete_dof", [n, 7 + nDOF])
    (Tx Ty Tz Qw Qx Qy Qz then DOFs; the caller composes mRefTran, planner.complete_dof_to_raw)."""
    if not states:
        return "aa_eigen", np.zeros((0, 0), np.float32)
    kinds = {(s.position_type, s.posture_type) for s in states}
    if len(kinds) != 1:
        raise ValueError("mixed state layouts in one batch")
    (st, pt), = kinds
    arr = np.stack([np.concatenate([s.position, s.posture]) for s in states]).astype(np.float32)
    if (st, pt) == (SPACE_AXIS_ANGLE, POSE_EIGEN):
        return "aa_eigen", arr
    if (st, pt) == (SPACE_COMPLETE, POSE_DOF):
        return "complete_dof", arr
    raise ValueError(f"state layout (position {st}, posture {pt}) has no batched evaluation path")
