"""C++ planner host (graspit_b200/host/b200Planner.{h,cpp}: SimAnnParams presets, SearchVariable tables, best list,
ListPlanner rule, state files, batched annealing over the C ABI) through its harness tests/host/b200planner_check.cpp
(built by graspit_b200.build.build_host_planner_check)."""
import os
import subprocess

import numpy as np
import pytest

from graspit_b200.formats.scenepack import load_pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "graspit_b200", "build", "b200planner_check")


def _dump_scene(tmp_path):
    """the scene as plain numbers (format: tests/host/b200planner_check.cpp::loadScene)"""
    sc = load_pack("barrett_mug")
    r = sc.robots[0]
    obj = [i for i, b in enumerate(sc.bodies) if b.kind == "object"][0]
    f = lambda xs: " ".join(repr(float(x)) for x in xs)
    lines = [str(len(sc.bodies))]
    for i, b in enumerate(sc.bodies):
        p = str(tmp_path / f"body{i}.tri")
        t = np.ascontiguousarray(b.tris, dtype=np.float32).reshape(-1, 9)
        with open(p, "wb") as fh:
            fh.write(np.int32(len(t)).tobytes()); fh.write(t.tobytes())
        q, tr = sc.body_tran[i]
        lines.append(f"{f(q)} {f(tr)} {p}")
    dis = list(sc.disabled_pairs) + list(sc.touching_pairs)
    lines.append(str(len(dis)) + " " + " ".join(f"{a} {b}" for a, b in dis))
    lines.append(f"{r.base_body} {r.mount_body} {r.n_dof}")
    lines.append(f(r.dof_min)); lines.append(f(r.dof_max))
    joints, links, last = [], [], []
    lines.append(str(len(r.chains)))
    for c in r.chains:
        lines.append(f"{f(c.tran[0])} {f(c.tran[1])} {len(joints)} {len(c.joints)} {len(links)} {len(c.link_bodies)}")
        joints += c.joints; links += c.link_bodies; last += list(c.last_joint)
    lines.append(str(len(joints)))
    for j in joints:
        lines.append(f"{j.dof} {int(j.revolute)} {f([j.ratio, j.theta, j.d, j.a, j.alpha])}")
    lines.append(str(len(links)))
    lines.append(" ".join(str(x) for x in links)); lines.append(" ".join(str(x) for x in last))
    lines.append(f"{f(r.approach[0])} {f(r.approach[1])}")
    lines.append(str(r.eg.shape[0])); lines.append(f(r.eg.reshape(-1))); lines.append(f(r.eg_origin)); lines.append(f(r.eg_norm))
    lines.append(str(len(r.vcontacts)))
    for v in r.vcontacts:
        lines.append(f"{v.body} {f(v.loc)} {f(v.normal)}")
    lines.append(str(obj)); lines.append(f"{f(sc.body_tran[obj][0])} {f(sc.body_tran[obj][1])}")
    # DOF data of Hand::autoGrasp: default velocities, break-away flags
    lines.append(f(r.dof_velocity)); lines.append(" ".join("1" if t == "b" else "0" for t in r.dof_type))
    p = str(tmp_path / "scene.txt")
    open(p, "w").write("\n".join(lines) + "\n")
    return p


def test_host_planner_logic_cpu():
    assert os.path.exists(BIN), "run `python __graft_entry__.py` (graspit_b200.build.build_host_planner_check)"
    r = subprocess.run([BIN, "cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]


@pytest.mark.gpu
def test_host_planner_batched_annealing_gpu(tmp_path):
    r = subprocess.run([BIN, "gpu", _dump_scene(tmp_path)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
