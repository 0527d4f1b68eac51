"""CPU: the autograsp control flow pinned to the reference's OWN lines.

oracle/ref_autograsp.cpp compiles Hand::autoGrasp, Robot::moveDOFToContacts / jumpDOFToContact / getJointValuesFromDOF /
interpolateJoints / stopJointsFromLink, the DOF classes' accumulateMove / updateVal / getJointValues and the pair filter of
World::findContacts -- cut out of the reference's sources at build time -- against stand-in Robot / World classes on the
reference's own collision backend.  The restatement the GPU tests compare the engine with (oracle/port/autograsp.py) has
to close every hand exactly as those lines do: same DOF and joint values, same stopped joints, same number of
moveDOFToContacts cycles, same outcome."""
import copy

import numpy as np
import pytest

from graspit_b200.formats.scenepack import load_pack
from graspit_b200.sampling import object_body, sample_aa_states
from oracle import ref
from oracle.port.autograsp import AutoGraspOracle, with_static_ratios
from tests.common import OracleScene, aa_to_raw_states

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def _open_hand_states(sc, n, seed):
    r = sc.robots[0]
    pos, _ = sample_aa_states(sc, n, seed=seed, strata=("close",))
    rest = np.clip(r.dof_default, r.dof_min, r.dof_max)
    return aa_to_raw_states(sc, pos, np.tile(rest[None, :], (n, 1)).astype(np.float32))


def _legal_starts(sc, osc, raw, want):
    """states whose open hand is collision free (what the planner hands to autoGrasp)"""
    legal = osc.eval(0, object_body(sc), raw)[0]
    return [int(p) for p in np.nonzero(np.asarray(legal))[0][:want]]


@pytest.mark.parametrize("types,stop_at_contact,nposes", [(None, False, 10), (None, True, 6), (["r", "c", "c", "c"], False, 6),
                                                         (["r", "r", "r", "r"], False, 4)])
def test_port_closes_hands_like_the_reference_lines(types, stop_at_contact, nposes):
    sc = copy.deepcopy(load_pack("barrett_mug"))
    if types is not None:
        sc.robots[0].dof_type = list(types)
    R = ref.RefAutoGrasp(sc, 0)                      # coupling ratios in; the reference's getStaticRatio lines do the rest
    # what the engine and the restatement are given: DOF::getStaticRatio per joint (CompliantDOF's differs from the coupling
    # ratio, dof.cpp:779-795) and the DOF limits that follow from them
    sc = with_static_ratios(sc, 0)
    # the reference's DOF::updateMinMax on the joint limits gives the same DOF limits
    assert np.allclose(R.dof_min, sc.robots[0].dof_min) and np.allclose(R.dof_max, sc.robots[0].dof_max)
    osc = OracleScene(sc)
    raw = _open_hand_states(sc, 400, 71)
    sel = _legal_starts(sc, osc, raw, nposes)
    assert len(sel) == nposes
    port = AutoGraspOracle(sc, 0, osc.w, osc.ids)
    moved_any = stopped_any = 0
    for p in sel:
        base = (raw[p, 0:4].astype(np.float64), raw[p, 4:7].astype(np.float64))
        dof0 = raw[p, 7:].astype(np.float64)
        a = R.auto_grasp(base, dof0, 1.0, stop_at_contact)
        b = port.auto_grasp(base, dof0, 1.0, stop_at_contact)
        assert a["moved"] == b["moved"] and a["error"] == b["error"], (p, a, b)
        assert not a["failsafe"]
        assert np.abs(a["dof"] - b["dof"]).max() < 1e-12, (p, a["dof"], b["dof"])
        assert np.abs(a["joints"] - b["joints"]).max() < 1e-12, (p, a["joints"], b["joints"])
        # cycles: the port counts the last one (the cycle that finds nothing left to move, or fails) as well
        ended_by_stop = stop_at_contact and a["moved"] and b["iters"] == a["jumps"]
        assert b["iters"] == a["jumps"] + (0 if ended_by_stop else 1), (p, a["jumps"], b["iters"])
        if not ended_by_stop and not a["error"]:
            # stoppedJoints as the DOFs last saw it = its final state when the loop ends on "no movement"
            assert (a["stopped"] == b["stopped"]).all(), (p, a["stopped"], b["stopped"])
        moved_any += a["moved"]
        stopped_any += int((a["stopped"] != 0).any())
    assert moved_any >= nposes - 1 and (stop_at_contact or stopped_any >= 1)
    R.close()


def test_loader_static_ratios_of_a_compliant_hand_match_the_reference_lines():
    """models/robots/harvardHand (four <dof type="c">, springStiffness per joint): the loader's JointDesc.ratio
    (CompliantDOF::getStaticRatio) and DOF limits against DOF::initDOF / updateMinMax / getStaticRatio of the reference,
    run on the coupling ratios and stiffnesses of the same file."""
    import os
    from graspit_b200.formats import xmlmodel
    from graspit_b200.formats.xmlmodel import Scene
    path = "/root/reference/models/robots/harvardHand/harvardHand.xml"
    if not os.path.exists(path):
        pytest.skip("reference model files not present (GPU box)")
    sc = Scene()
    ridx = xmlmodel.load_robot(sc, path, (np.array([1.0, 0, 0, 0]), np.zeros(3)))
    r = sc.robots[ridx]
    assert "c" in r.dof_type
    joints = [j for c in r.chains for j in c.joints]
    changed = [j for j in joints if j.coupling is not None]
    assert len(changed) == 4          # the odd-numbered joints went through the curvature formula (here it lands on 0.6 again)
    if r.dof_velocity is None:
        r.dof_velocity = np.ones(r.n_dof)
    R = ref.RefAutoGrasp(sc, ridx)
    assert np.allclose(R.dof_min, r.dof_min, rtol=1e-12, atol=1e-12), (R.dof_min, r.dof_min)
    assert np.allclose(R.dof_max, r.dof_max, rtol=1e-12, atol=1e-12), (R.dof_max, r.dof_max)
    # joint values the reference's DOFs hand out at the DOF limits = ratio x limit of the loader
    base = (np.array([1.0, 0, 0, 0]), np.zeros(3))
    a = R.auto_grasp(base, r.dof_max, 0.0, False)          # speed 0: no motion, but the posture has been forced
    want = np.array([j.ratio * r.dof_max[j.dof] for j in joints])
    assert np.allclose(a["joints"], want, rtol=1e-12, atol=1e-12)
    R.close()
