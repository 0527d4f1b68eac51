"""A11 / A12 pinned: the numpy restatement of forward kinematics and of the state -> pose mapping (oracle/port/fk.py, the
oracle of every FK parity test) against the reference's OWN LINES -- DHTransform, Revolute/PrismaticJoint::getTran,
KinematicChain::fwdKinematics and the four PositionState*::getCoreTran bodies, cut out of /root/reference at build time
and compiled against the reference's transf (oracle/ref_fk.cpp, oracle/Makefile)."""
import numpy as np
import pytest

from graspit_b200.formats.scenepack import load_pack
from graspit_b200.sampling import object_body
from oracle import ref
from oracle.port import fk as ofk

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle library not built")


def _same(a, b, tq=1e-13, tt=1e-11):
    assert min(np.abs(a[0] - b[0]).max(), np.abs(a[0] + b[0]).max()) < tq, (a[0], b[0])
    assert np.abs(a[1] - b[1]).max() < tt, (a[1], b[1])


@pytest.mark.parametrize("pack", ["barrett_mug", "humanhand_flask", "staubli_barrett", "dlr_flask"])
def test_port_fk_equals_reference_lines(pack):
    sc = load_pack(pack)
    rng = np.random.default_rng(11)
    root = 0
    n = 12
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.uniform(-300, 300, size=(n, 3))
    dofs = {k: rng.uniform(r.dof_min, r.dof_max, size=(n, r.n_dof)) for k, r in enumerate(sc.robots)}
    # small joint angles exercise the AXIS_ANGLE_ROTATION identity snap (transform.cpp:26-29)
    for k in dofs:
        dofs[k][0] = 0.0
        dofs[k][1] = 1.0e-4
    port = ofk.forward_kinematics(sc.robots, root, (q, t), dofs)
    for p in range(n):
        got = ref.forward_kinematics(sc.robots, root, (q[p], t[p]), {k: v[p] for k, v in dofs.items()})
        assert set(got) == set(port)
        for b in got:
            _same((port[b][0][p], port[b][1][p]), got[b])


def test_port_state_mappings_equal_reference_lines():
    sc = load_pack("barrett_mug")
    r = sc.robots[0]
    ref_tran = sc.body_tran[object_body(sc)]
    rng = np.random.default_rng(3)
    n = 64
    aa = np.stack([rng.uniform(-250, 250, n), rng.uniform(-250, 250, n), rng.uniform(-250, 350, n), rng.uniform(0, np.pi, n),
                   rng.uniform(-np.pi, np.pi, n), rng.uniform(0, np.pi, n)], 1)
    aa[0, 5] = 0.0; aa[1, 5] = 5.0e-4          # alpha below the identity snap
    port = ofk.aa_state_to_base(r, ref_tran, aa)
    for p in range(n):
        _same((port[0][p], port[1][p]), ref.state_to_base(1, aa[p], ref_tran, r.approach))
    comp = np.concatenate([rng.uniform(-250, 250, (n, 3)), rng.uniform(-5, 5, (n, 4))], 1)
    port = ofk.complete_state_to_base(ref_tran, comp)
    for p in range(n):
        _same((port[0][p], port[1][p]), ref.state_to_base(0, comp[p], ref_tran, r.approach))
    ell = np.stack([rng.uniform(-np.pi / 2, np.pi / 2, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(-50, 100, n)], 1)
    ell[0, 2] = 0.0
    port = ofk.ellipsoid_state_to_base(r, ref_tran, ell)
    for p in range(n):
        _same((port[0][p], port[1][p]), ref.state_to_base(2, ell[p], ref_tran, r.approach), tq=1e-12, tt=1e-10)
    port = ofk.ellipsoid_state_to_base(r, ref_tran, ell, abc=(60.0, 90.0, 120.0))        # a != b: the normal is not unit length
    for p in range(n):
        _same((port[0][p], port[1][p]), ref.state_to_base(2, ell[p], ref_tran, r.approach, params=(60.0, 90.0, 120.0)), tq=1e-12, tt=1e-10)
    app = np.stack([rng.uniform(-30, 200, n), rng.uniform(-np.pi / 3, np.pi / 3, n), rng.uniform(-np.pi / 3, np.pi / 3, n)], 1)
    app[0, 1] = 0.0; app[1, 2] = 2.0e-4
    port = ofk.approach_state_to_base(r, ref_tran, app)
    for p in range(n):
        _same((port[0][p], port[1][p]), ref.state_to_base(3, app[p], ref_tran, r.approach))
