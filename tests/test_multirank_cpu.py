"""N > 1 host logic on CPU: world_size-2 gloo.  The path shards by pose / chain with no data-path collective; the
only exchange is the all-gather of each rank's best (state || energy) rows for the planner's best list."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graspit_b200.planner import BEST_LIST_SIZE, BestList, exchange_best, local_best_rows, shard

REF = (np.array([1.0, 0, 0, 0]), np.zeros(3))


def _candidates(total, nvar=8, seed=0):
    rng = np.random.default_rng(seed)
    st = rng.uniform(-100, 100, size=(total, nvar)).astype(np.float32)
    st[:, 3] = rng.uniform(0, np.pi, total); st[:, 4] = rng.uniform(-np.pi, np.pi, total); st[:, 5] = rng.uniform(0, np.pi, total)
    en = rng.uniform(5, 80, size=total).astype(np.float32)
    en[rng.random(total) < 0.3] = 1.0e8           # chains that never found a legal state
    return st, en


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st, en = _candidates(total)
    lo, hi = shard(total, rank, world)
    rows = local_best_rows(torch.from_numpy(st[lo:hi]), torch.from_numpy(en[lo:hi]), BEST_LIST_SIZE)
    allrows = exchange_best(rows)
    assert allrows.shape == (world * BEST_LIST_SIZE, st.shape[1] + 1)
    # every rank sees the same gathered rows, rank-major
    mine = allrows[rank * BEST_LIST_SIZE:(rank + 1) * BEST_LIST_SIZE]
    assert torch.equal(mine, rows)
    bl = BestList(REF)
    a = allrows.numpy()
    bl.merge(a[:, :-1], a[:, -1])
    np.save(os.path.join(out_dir, f"best_{rank}.npy"), np.array(bl.energies))
    dist.barrier()
    dist.destroy_process_group()


def test_best_list_exchange_world2_gloo(tmp_path):
    total, world = 5000, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    e0, e1 = np.load(tmp_path / "best_0.npy"), np.load(tmp_path / "best_1.npy")
    assert (e0 == e1).all()
    # equals the list a single process builds from the per-shard top-20 candidates
    st, en = _candidates(total)
    cand_s, cand_e = [], []
    for r in range(world):
        lo, hi = shard(total, r, world)
        idx = np.argsort(en[lo:hi], kind="stable")[:BEST_LIST_SIZE] + lo
        cand_s.append(st[idx]); cand_e.append(en[idx])
    bl = BestList(REF)
    bl.merge(np.concatenate(cand_s), np.concatenate(cand_e))
    assert np.allclose(np.array(bl.energies), e0)
    # and its head equals the global optimum over all chains
    assert abs(e0[0] - en.min()) < 1e-6
    assert len(e0) <= BEST_LIST_SIZE and (np.diff(e0) >= 0).all()
