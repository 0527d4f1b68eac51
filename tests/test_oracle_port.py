"""The numpy restatement (oracle/port) against the compiled reference transf and against load-time poses."""
import numpy as np
import pytest

from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import fk as ofk
from oracle.port import transf_np as T

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def test_transf_port_matches_reference():
    rng = np.random.default_rng(0)
    for _ in range(50):
        qa, qb = rng.normal(size=4), rng.normal(size=4)
        ta, tb = rng.uniform(-100, 100, 3), rng.uniform(-100, 100, 3)
        r = ref.transf_compose((qa, ta), (qb, tb))
        a, b = T.make(qa[None], ta[None]), T.make(qb[None], tb[None])
        m = T.compose(a, b)
        assert np.allclose(m[0][0], r[0], atol=1e-14) and np.allclose(m[1][0], r[1], atol=1e-11)
        ri = ref.transf_inverse((qa, ta))
        mi = T.inverse(a)
        assert np.allclose(mi[0][0], ri[0], atol=1e-14) and np.allclose(mi[1][0], ri[1], atol=1e-11)
        p = rng.uniform(-50, 50, 3)
        assert np.allclose(T.apply(a, p[None])[0], ref.transf_apply((qa, ta), p), atol=1e-11)
        ang = rng.uniform(-4, 4); ax = rng.normal(size=3)
        rq, _ = ref.transf_axis_angle(ang, ax)
        mq, _ = T.axis_angle_rotation(np.array([ang]), ax)
        assert np.allclose(mq[0], rq, atol=1e-14)
    # the identity snap
    assert np.array_equal(T.axis_angle_rotation(np.array([5e-4]), [0, 0, 1])[0][0], [1, 0, 0, 0])


def test_fk_zero_posture_matches_dh_by_hand():
    sc = load_pack("barrett_mug")
    r = sc.robots[0]
    out = ofk.forward_kinematics(sc.robots, 0, T.identity(1), {0: np.zeros((1, 4))})
    # chain 0 link 0: base % chain.tran % Rz(0 + 90deg) % Tx(50) % Rx(-90deg)
    assert out[2][1][0] == pytest.approx([25.0, 50.0, -1.0], abs=1e-9)
    # mirror chain
    assert out[5][1][0] == pytest.approx([-25.0, 50.0, -1.0], abs=1e-9)
    # thumb chain has two links only
    assert set(out.keys()) == set(range(1, 10))
    # coupling ratio 1/3 on the distal joints
    out2 = ofk.forward_kinematics(sc.robots, 0, T.identity(1), {0: np.array([[0.0, 0.9, 0.0, 0.0]])})
    assert not np.allclose(out2[4][1], out[4][1])
    assert np.allclose(out2[7][1], out[7][1])


def test_scene_pack_load_poses_are_fk_of_world_file():
    sc = load_pack("barrett_mug")
    r = sc.robots[0]
    base = (np.asarray(r.tran[0])[None], np.asarray(r.tran[1])[None])
    out = ofk.forward_kinematics(sc.robots, 0, base, {0: np.asarray(r.dof_values)[None]})
    for b, (q, t) in out.items():
        assert np.allclose(sc.body_tran[b][0], q[0], atol=1e-12)
        assert np.allclose(sc.body_tran[b][1], t[0], atol=1e-9)


def test_eigen_to_dof_clamps():
    sc = load_pack("barrett_mug")
    r = sc.robots[0]
    d = ofk.eigen_to_dof(r, np.array([[100.0, 100.0], [-100.0, -100.0], [0.0, 0.0]]))
    assert np.allclose(d[0], r.dof_max) and np.allclose(d[1], r.dof_min)
    assert np.allclose(d[2], r.eg_origin)
