"""Planner-state text files (HandObjectState::writeToFile / readFromFile wire format) and the batched ListPlanner rule."""
import io

import numpy as np

from graspit_b200.formats import statefile as sf
from graspit_b200.planner import complete_dof_to_raw


def test_format_is_the_reference_printf_format():
    s = sf.PlannerState(sf.POSE_EIGEN, np.array([0.5, -1.25]), sf.SPACE_AXIS_ANGLE, np.array([10.0, -20.5, 300.0, 0.1, -3.0, 1.5707963]))
    buf = io.StringIO()
    sf.write_states(buf, [s])
    # fprintf(fp, "%d ", type); fprintf(fp, "%f ", value) ...; fprintf(fp, "\n")   (searchState.cpp:224-232)
    assert buf.getvalue() == "4 0.500000 -1.250000 \n1 10.000000 -20.500000 300.000000 0.100000 -3.000000 1.570796 \n"


def test_round_trip_and_truncated_input():
    rng = np.random.default_rng(0)
    states = [sf.PlannerState(sf.POSE_EIGEN, rng.uniform(-4, 4, 2), sf.SPACE_AXIS_ANGLE, rng.uniform(-3, 3, 6)) for _ in range(50)]
    buf = io.StringIO()
    sf.write_states(buf, states)
    back = sf.read_states(io.StringIO(buf.getvalue()), n_posture=2)
    assert len(back) == 50
    for a, b in zip(states, back):
        assert a.posture_type == b.posture_type and a.position_type == b.position_type
        assert np.allclose(a.posture, b.posture, atol=5e-7) and np.allclose(a.position, b.position, atol=5e-7)
    kind, arr = sf.to_engine_batch(back)
    assert kind == "aa_eigen" and arr.shape == (50, 8) and arr.dtype == np.float32
    assert np.allclose(arr[:, :6], np.stack([s.position for s in states]), atol=1e-5)
    # a state whose position line is too short ends the read (reference: "Line to short to read all state variables")
    text = buf.getvalue().splitlines(True)
    text[7] = "1 1.0 2.0 \n"
    assert len(sf.read_states(io.StringIO("".join(text)), n_posture=2)) == 3
    # wrong type codes are refused
    assert sf.read_states(io.StringIO("7 1.0 \n1 0 0 0 0 0 0 \n")) == []
    assert sf.read_states(io.StringIO("")) == []


def test_complete_dof_layout():
    st = [sf.PlannerState(sf.POSE_DOF, np.array([0.1, 0.2, 0.3, 0.4]), sf.SPACE_COMPLETE, np.array([1.0, 2.0, 3.0, 1.0, 0.0, 0.0, 0.0]))]
    kind, arr = sf.to_engine_batch(st)
    assert kind == "complete_dof" and arr.shape == (1, 11)
    # base = mRefTran % transf(Q, T): a reference frame rotated 90 deg about z, translated by (10, 0, 0)
    c = np.sqrt(0.5)
    raw = complete_dof_to_raw(arr, (np.array([c, 0, 0, c]), np.array([10.0, 0, 0])))
    assert np.allclose(raw[0, :4], [c, 0, 0, c], atol=1e-6)
    assert np.allclose(raw[0, 4:7], [10.0 - 2.0, 1.0, 3.0], atol=1e-5)
    assert np.allclose(raw[0, 7:], [0.1, 0.2, 0.3, 0.4], atol=1e-6)
