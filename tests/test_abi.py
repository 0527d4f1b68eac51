"""The C-ABI library loads and exports every symbol include/b2c.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from graspit_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b2c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2c_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libb2c.so does not export {n}"
    assert set(engine.EXPORTS) == set(names), set(engine.EXPORTS) ^ set(names)


def test_no_cpu_fallback_create_fails_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = engine.load_library()
    h = ctypes.c_void_p()
    rc = lib.b2c_create(0, ctypes.byref(h))
    assert rc == -4 and not h          # B2C_ENODEVICE
    with pytest.raises(engine.B2CError):
        engine.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "graspit_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".inc")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_host_obb_hierarchy_matches_reference_goldens():
    # getBoundingVolumes parity (test/graspit_collision_tests.cpp:90-140) through the host-only entry point
    from graspit_b200.formats.scenepack import load_pack
    sc = load_pack("barrett_mug")
    bm = engine.obb_hierarchy(sc.bodies[0].tris, 0)
    assert bm.shape == (1, 10)
    assert bm[0, 0:3] == pytest.approx([-11.662, -8.98983, 0.000190189], abs=1e-3)
    assert bm[0, 3:7] == pytest.approx([0.999999, -9.15765e-06, -1.47701e-06, 0.00157698], abs=1e-3)
    bp = engine.obb_hierarchy(sc.bodies[1].tris, 0)
    assert bp[0, 0:3] == pytest.approx([0.013166, -9.91786, -38.5818], abs=1e-3)
    assert bp[0, 3:7] == pytest.approx([0.6664499, 0.236398, -0.66585779, -0.23789390], abs=1e-3)


def test_host_obb_hierarchy_matches_oracle_at_every_depth():
    from graspit_b200.formats.scenepack import load_pack
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    sc = load_pack("barrett_mug")
    w = ref.RefWorld()
    for bi in (0, 1, 4):
        oid = w.add_body(sc.bodies[bi].tris.astype(np.float64))
        for depth in (0, 2, 5, 40):
            a = engine.obb_hierarchy(sc.bodies[bi].tris, depth)
            b = w.bounding_volumes(oid, depth)
            assert a.shape == b.shape
            assert np.abs(a[:, :3] - b[:, :3]).max() < 1e-9 and np.abs(a[:, 7:] - b[:, 7:]).max() < 1e-9
            sgn = np.sign((a[:, 3:7] * b[:, 3:7]).sum(1, keepdims=True))
            assert np.abs(a[:, 3:7] - sgn * b[:, 3:7]).max() < 1e-9


def test_contact_postprocess_matches_reference():
    """removeContactDuplicates + compactContactSet (collisionInterface.cpp:74-287) on the reference's own raw
    candidate lists, including a flat patch that exercises the 2-D perimeter hull."""
    from graspit_b200.formats.scenepack import load_pack
    from oracle import ref
    from oracle.port import transf_np as T
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    sc = load_pack("barrett_mug")
    w = ref.RefWorld()
    mug = w.add_body(sc.bodies[0].tris.astype(np.float64))
    palm = w.add_body(sc.bodies[1].tris.astype(np.float64))
    # a 60 x 60 mm plate, finely triangulated, to press the palm's flat face on
    g = np.linspace(-30, 30, 7)
    tris = []
    for i in range(6):
        for j in range(6):
            a, b, c, d = [g[i], g[j], 0], [g[i + 1], g[j], 0], [g[i + 1], g[j + 1], 0], [g[i], g[j + 1], 0]
            tris += [[a, b, c], [a, c, d]]
    plate = w.add_body(np.array(tris, dtype=np.float64))
    rng = np.random.default_rng(7)
    checked = 0
    key = lambda rows: sorted(tuple(np.round(r, 9)) for r in rows)
    for trial in range(40):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.uniform(-150, 150, 3) + np.array([0, 0, 105.0])
        w.set_transform(mug, [1, 0, 0, 0], [0, 0, 0]); w.set_transform(palm, q, t)
        d, p1, p2 = w.body_distance(palm, mug)
        if d <= 0:
            continue
        tr = T.make(q[None], t[None])
        pw1 = T.apply(tr, T.apply(tr, p1[None]))[0]          # undo the reference's double inverse
        dirv = (p2 - pw1) / np.linalg.norm(p2 - pw1)
        w.set_transform(palm, q, t + dirv * (d - 0.1))
        raw, _ = w.contact_raw(palm, mug, 0.2)
        full = w.contact(palm, mug, 0.2)
        mine = engine.contacts_postprocess(raw, 3.0, True)
        assert key(full) == key(mine)
        checked += 1
    assert checked >= 20
    # flat-on-flat: a coarse 40 x 30 mm plate 0.1 mm above the fine one -> coplanar contact patch
    top = w.add_body(np.array([[[-20, -15, 0], [20, -15, 0], [20, 15, 0]], [[-20, -15, 0], [20, 15, 0], [-20, 15, 0]]],
                              dtype=np.float64))
    w.set_transform(top, ref.transf_axis_angle(0.3, [0, 0, 1])[0], [1.5, -2.0, 0.1])
    w.set_transform(plate, [1, 0, 0, 0], [0, 0, 0])
    raw, _ = w.contact_raw(top, plate, 0.2)
    full = w.contact(top, plate, 0.2)
    mine = engine.contacts_postprocess(raw, 3.0, True)
    assert len(raw) > 10 and len(full) >= 3
    assert key(full) == key(mine)
