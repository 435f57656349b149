"""CPU suite: the N>1 exchange (best-key min-allreduce + elite broadcast) on gloo, world 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from constraint_solver_b200.dist import exchange_best, exchange_best_device, owner_of


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cpr, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    rows = rng.integers(0, n, size=(cpr, n)).astype(np.int16)  # this rank's chain shard
    scores = rng.integers(5, 1000, size=cpr)
    if rank == 1:
        scores[3] = 2  # global best lives on rank 1, local chain 3
    best_local = int(scores.argmin())
    key = torch.tensor([(int(scores[best_local]) << 32) | (rank * cpr + best_local)], dtype=torch.int64)
    elite = torch.zeros(n, dtype=torch.int16)
    score, gid = exchange_best(dist, key, lambda c: torch.from_numpy(rows[c]), elite, rank, cpr)
    ok = (score == 2 and gid == 1 * cpr + 3)
    # every rank must now hold rank 1's chain 3
    expect = np.random.default_rng(101).integers(0, n, size=(cpr, n)).astype(np.int16)[3]
    ok = ok and np.array_equal(elite.numpy(), expect)
    # the host-sync-free variant delivers the same key and the same elite
    stride = n + 3
    rows_pad = torch.zeros((cpr, stride), dtype=torch.int16)
    rows_pad[:, :n] = torch.from_numpy(rows)
    elite2 = torch.full(((n + 1) & ~1,), 7, dtype=torch.int16)
    key2 = exchange_best_device(dist, key, rows_pad, elite2, rank, cpr)
    ok = ok and int(key2.item()) == ((2 << 32) | (1 * cpr + 3))
    ok = ok and np.array_equal(elite2.numpy()[:n], expect) and int(elite2[n:].abs().sum()) == 0
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_owner_of_roundtrip():
    assert owner_of((7 << 32) | 9000, 4096) == (7, 2, 808)
    assert owner_of(5, 4096) == (0, 0, 5)


@pytest.mark.timeout(120)
def test_best_exchange_world2_gloo():
    world, cpr, n = 2, 8, 33
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), cpr, n, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
