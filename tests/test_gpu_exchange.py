"""The collective-free best-row exchange (b2c_exchange_* / b2c_anneal_publish): two contexts in one process, each with its own
chains, publish their BEST_LIST_SIZE best (state || energy) rows into each other's ring; collect returns the rows of both
ranks, equal to what each annealer's own best list holds (b2c_anneal_best_rows), tagged with the step they belong to.
Runs on one GPU (both contexts on device 0); with two GPUs the second context moves to device 1 (peer access over NVLink)."""
import numpy as np
import pytest
import torch

from graspit_b200.engine import Annealer, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from graspit_b200.planner import BEST_LIST_SIZE, BestList, PeerExchange, SearchVariables, SimAnnParams
from graspit_b200.sampling import object_body

pytestmark = pytest.mark.gpu


def _rank(dev, rank, nchain, scene):
    eng = Engine(dev)
    sb = SceneBinding(eng, scene)
    obj = object_body(scene)
    sv = SearchVariables.axis_angle_eigen(scene.robots[0].eg.shape[0])
    rng = np.random.default_rng(100 + rank)
    init = rng.uniform(sv.vmin, sv.vmax, size=(nchain, len(sv.vmin))).astype(np.float32)
    an = Annealer(eng, sb.robot_ids[0], sb.body_ids[obj], sv, SimAnnParams.ANNEAL_DEFAULT(), init, scene.body_tran[obj], seed=7,
                  chain_offset=rank * nchain)
    return eng, an


def test_peer_exchange_two_contexts():
    scene = load_pack("barrett_mug")
    ndev = torch.cuda.device_count()
    devs = [0, 1 if ndev > 1 else 0]
    nchain = 2048
    ranks = [_rank(devs[r], r, nchain, scene) for r in range(2)]
    xs = [PeerExchange(eng, an, r, 2, rows=BEST_LIST_SIZE, depth=4, peers=True) for r, (eng, an) in enumerate(ranks)]
    PeerExchange.connect_in_process(xs)
    nvar = ranks[0][1].nvar
    for step in range(6):
        own = []
        for r, (eng, an) in enumerate(ranks):
            an.step(1)
            xs[r].publish()
        for r, (eng, an) in enumerate(ranks):
            eng.check_overflow()                         # synchronises the context's stream: both publishes have landed
        for r, (eng, an) in enumerate(ranks):
            rows = torch.empty((BEST_LIST_SIZE, nvar + 1), dtype=torch.float32, device=f"cuda:{devs[r]}")
            an.best_rows(BEST_LIST_SIZE, rows.data_ptr())
            eng.check_overflow()
            own.append(rows.cpu().numpy())
        for r in range(2):
            got, steps = xs[r].collect()
            assert steps.tolist() == [step + 2, step + 2]                 # tag = annealing steps completed + 1; newest complete slot of either source
            for src in range(2):
                assert np.array_equal(got[src].numpy(), own[src]), (step, r, src)
    # rows are real: finite energies of legal states, sorted into a best list without error
    bl = BestList(scene.body_tran[object_body(scene)])
    xs[0].merge_into(bl)
    assert 1 <= len(bl.states) <= BEST_LIST_SIZE and np.isfinite(bl.energies).all()
    assert bl.energies == sorted(bl.energies)
    for eng, an in ranks:
        an.close()
        eng.exchange_destroy()


def test_collect_before_anything_arrived_reports_step_zero():
    scene = load_pack("barrett_mug")
    eng, an = _rank(0, 0, 256, scene)
    x = PeerExchange(eng, an, 0, 1, rows=BEST_LIST_SIZE)
    rows, steps = x.collect()
    assert steps.tolist() == [0] and (rows.numpy()[..., -1] > 1e37).all()
    an.step(1)
    x.publish()
    rows, steps = x.collect()
    assert steps.tolist() == [2]          # one step completed
