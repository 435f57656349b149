"""GPU suite for the host-sync-free multi-GPU loop: cs_*_step_enqueue / cs_*_step_wait and the
owner-masked elite gather cs_*_exchange_select, plus BestExchange.sync_device() on a
one-rank NCCL group against the torch-op formulation the gloo CPU test pins."""
import os
import socket

import numpy as np
import pytest
import torch

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _nq(offset=0, chains=6, n=96):
    e = cs.NQueensChains(n, chains, seed=5, chain_offset=offset)
    e.init_random()
    return e


def _es(offset=0, chains=5):
    ids = np.array([3, 8, 9, 20, 41], dtype=np.int64)
    e = cs.ScheduleChains(21, ids, holidays=[(8, 2), (41, 20)], n_chains=chains, seed=9, chain_offset=offset)
    e.init_random()
    return e


@pytest.mark.parametrize("make", [_nq, _es])
def test_enqueue_plus_wait_is_step(make):
    a, b = make(), make()
    for k in (1, 3, 64):
        sa = a.step(k)
        b.step_enqueue(k)
        sb = b.step_wait()
        assert sa.moves_scored == sb.moves_scored and sa.steps_accepted == sb.steps_accepted
        assert sa.best_chain == sb.best_chain
        assert np.array_equal(a.get_chains(), b.get_chains())
    # several launches in flight, one wait: the stream orders them
    for _ in range(4):
        b.step_enqueue(2)
    b.step_wait()
    a.step(8)
    assert np.array_equal(a.get_chains(), b.get_chains())
    with pytest.raises(cs.CsError) as ei:
        b.step_wait()
    assert ei.value.status == L.CS_ERR_STATE
    a.close(), b.close()


def _select(eng, key_value, elite_len):
    key = torch.tensor([key_value], dtype=torch.int64, device="cuda")
    elite = torch.full((elite_len,), 0x5A5A, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    eng.exchange_select(key.data_ptr(), elite.data_ptr(), elite_len)
    eng.scores()   # a call that waits for the handle's stream
    return elite.cpu().numpy().astype(np.int64) & 0xFFFF


def test_exchange_select_nqueens_owner_and_stranger():
    off, chains, n = 4096, 6, 96
    e = _nq(off, chains, n)
    e.step(2)
    rows = e.get_chains()
    for local in (0, 3, chains - 1):
        got = _select(e, (17 << 32) | (off + local), n + 2)
        assert np.array_equal(got[:n], rows[local]) and not got[n:].any()
    for gid in (off - 1, off + chains, 0, 0xFFFFFFFF):   # another rank's chain: zeros only
        assert not _select(e, (1 << 32) | gid, n + 2).any()
    assert np.array_equal(_select(e, off + 1, 10), rows[1][:10])   # short elite buffer: a prefix
    e.close()


def test_exchange_select_scheduling_owner_and_stranger():
    off, chains = 10, 5
    e = _es(off, chains)
    e.step(1)
    ids = np.array([3, 8, 9, 20, 41])
    rows = e.get_chains()          # employee ids, n_slots (+ phantom) per chain
    for local in range(chains):
        got = _select(e, (3 << 48) | (2 << 32) | (off + local), 24)
        T = e.n_slots                                              # scored slots + the phantom one
        assert np.array_equal(ids[got[:T]], rows[local])           # dense indices of the same rota
        assert not got[T:].any()
    assert not _select(e, off + chains, 24).any()
    assert not _select(e, off - 1, 24).any()
    e.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_best_exchange_one_rank_nccl_fused_equals_torch_ops():
    import torch.distributed as dist

    from constraint_solver_b200.dist import BestExchange

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        for eng in (_nq(0, 6, 96), _es(0, 5)):
            x = BestExchange(eng, dist, 0, 1, eng.n_chains)
            eng.step_enqueue(3)
            x.sync_device()
            fused_key, fused_elite = x.result(), x.elite.clone()
            st = eng.step_wait()
            x.sync_device_torch_ops()
            assert x.result() == fused_key
            assert torch.equal(x.elite, fused_elite)
            assert fused_key[1] == st.best_chain
            row = eng.get_chains()[st.best_chain]
            got = fused_elite.cpu().numpy().astype(np.int64) & 0xFFFF
            if isinstance(eng, cs.NQueensChains):
                assert np.array_equal(got[: eng.n], row)
            else:
                assert np.array_equal(np.array([3, 8, 9, 20, 41])[got[: eng.n_slots]], row)
            eng.close()
    finally:
        dist.destroy_process_group()
