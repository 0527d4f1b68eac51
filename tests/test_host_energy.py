"""The batched SearchEnergy / planner bindings (graspit_b200/host/b200ContactEnergy.cpp, b200SimAnnPlanner.cpp, compiled
against the reference's unmodified headers) A/B against the reference's own planner stack -- searchState.cpp,
searchStateImpl.cpp, egPlanner.cpp, simAnnPlanner.cpp, simAnn.cpp, simAnnParams.cpp compiled verbatim -- through the harness
tests/host/b200energy_check.cpp (built by oracle/Makefile into oracle/_ref/, where the reference objects live)."""
import os
import subprocess

import pytest

from tests.test_host_planner import _dump_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "b200energy_check")


def test_bindings_compile_against_reference_headers():
    """the harness links the product's SearchEnergy / EGPlanner subclasses with the verbatim-compiled reference planner objects"""
    if not os.path.isdir("/root/reference") and not os.path.exists(BIN):
        pytest.skip("reference sources and prebuilt harness both absent")
    assert os.path.exists(BIN), "run `make -C oracle` (needs /root/reference)"
    syms = subprocess.run(["nm", "-C", BIN], capture_output=True, text=True).stdout
    for s in ("B200ContactEnergy::analyzeStates", "B200SimAnnPlanner::mainLoop", "SimAnnPlanner::mainLoop", "SimAnn::iterate",
              "EGPlanner::addToListOfUniqueSolutions", "HandObjectState::HandObjectState", "PositionStateAA::getCoreTran"):
        assert s in syms, s


@pytest.mark.gpu
def test_search_energy_and_planner_bindings_gpu(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/b200energy_check not built")
    r = subprocess.run([BIN, _dump_scene(tmp_path)], capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
