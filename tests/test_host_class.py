"""Drop-in boundary: B200Collision (graspit_b200/host, C++ subclass of the reference's CollisionInterface) driven
side by side with the reference's GraspitCollision through the same CollisionInterface* by the C++ harness
tests/host/b200collision_check.cpp (built by oracle/Makefile into oracle/_ref/, next to the reference objects it links)."""
import os
import subprocess

import numpy as np
import pytest

from graspit_b200.formats.scenepack import load_pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "b200collision_check")


def _write_tri(path, tris):
    t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
    with open(path, "wb") as f:
        f.write(np.int32(len(t)).tobytes())
        f.write(t.tobytes())


def _files(tmp_path):
    sc = load_pack("barrett_mug")
    names = {"mug": 0, "palm": 1, "link": 2}
    out = []
    for n, i in names.items():
        p = str(tmp_path / f"{n}.tri")
        _write_tri(p, sc.bodies[i].tris)
        out.append(p)
    return out


def test_harness_present_and_refuses_to_run_without_a_device(tmp_path):
    """no CPU path: without a CUDA device the backend reports itself disabled and the harness exits 3"""
    assert os.path.exists(BIN), "run `make -C oracle` (needs /root/reference) to build the harness"
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([BIN] + _files(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_b200collision_matches_graspitcollision_through_the_interface(tmp_path):
    r = subprocess.run([BIN] + _files(tmp_path), capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout
