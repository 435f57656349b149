"""Multi-GPU plumbing: chains are sharded across ranks (one process per GPU, chain k of rank r
has global id r * chains_per_rank + k == its Philox stream), so the data path needs NO
collective.  The only exchange is the periodic best-score min-allreduce of the packed key
((score << 32) | global chain id) followed by the elite broadcast from the owning rank
(SURVEY 8e).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

import torch


class _DevArray:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (ptr, False), "version": 2}


def device_view(ptr: int, shape, typestr: str, device) -> torch.Tensor:
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def _bind_to_torch_stream(eng) -> None:
    """A handle's own stream is cudaStreamNonBlocking: it does not order against torch's current
    stream, where the collectives run.  Put the engine's kernels on torch's current stream so
    scan -> all_reduce -> apply (and step -> exchange) are ordered by the stream itself."""
    eng.set_stream(torch.cuda.current_stream().cuda_stream)


def owner_of(key: int, chains_per_rank: int):
    """(score, owning rank, local chain) of a packed best key."""
    gid = key & 0xFFFFFFFF
    return key >> 32, gid // chains_per_rank, gid % chains_per_rank


def exchange_best(dist, local_key: torch.Tensor, rows_of, elite: torch.Tensor, rank: int,
                  chains_per_rank: int):
    """min-allreduce the packed key, then broadcast the winner's solution into `elite`.

    local_key: 1-element int64 tensor (this rank's best key); rows_of(local_chain) -> tensor
    shaped like `elite` holding that chain's solution.  Returns (score, global chain id).
    """
    key = local_key.clone()
    dist.all_reduce(key, op=dist.ReduceOp.MIN)
    k = int(key.item())
    score, owner, local = owner_of(k, chains_per_rank)
    if owner == rank:
        elite.copy_(rows_of(local))
    dist.broadcast(elite.view(torch.uint8), src=owner)  # byte view: no int16 in NCCL/gloo
    return score, k & 0xFFFFFFFF


def exchange_best_device(dist, local_key: torch.Tensor, rows_all: torch.Tensor, elite: torch.Tensor,
                         rank: int, chains_per_rank: int) -> torch.Tensor:
    """The same exchange with NO host round trip: the reduced key is decoded
    on the device, the owning rank selects its chain with a device-side gather, every other rank
    contributes zeros, and one sum-all-reduce of the elite buffer delivers it everywhere.

    local_key: 1-element int64; rows_all: this rank's chains as an int16 [chains, stride] view
    (library-owned memory); elite: int16 buffer of even length >= solution length.  Everything
    is enqueued on the current stream; returns the reduced key as a DEVICE tensor.
    """
    key = local_key.clone()
    dist.all_reduce(key, op=dist.ReduceOp.MIN)
    gid = key & 0xFFFFFFFF
    owner = torch.div(gid, chains_per_rank, rounding_mode="floor")
    local = gid - owner * chains_per_rank                     # always a valid local index
    mine = (owner == rank).to(torch.int16)                    # 1 on the owning rank, else 0
    n = min(elite.numel(), rows_all.shape[1])
    elite.zero_()
    elite[:n] = rows_all.index_select(0, local)[0, :n] * mine
    dist.all_reduce(elite.view(torch.int32), op=dist.ReduceOp.SUM)  # zeros + the elite = the elite
    return key


class BestExchange:
    """exchange_best bound to an NQueensChains or ScheduleChains engine running on torch's
    current stream.  The packed key keeps the global chain id in its low 32 bits for both
    problems (n-queens: score << 32; scheduling: hard << 48 | soft << 32)."""

    def __init__(self, eng, dist, rank: int, world: int, chains_per_rank: int):
        self.eng, self.dist, self.rank, self.world = eng, dist, rank, world
        self.cpr = chains_per_rank
        _bind_to_torch_stream(eng)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.key = device_view(eng.best_key_device_ptr(), (1,), "<i8", dev)
        self.len = getattr(eng, "n", None) or eng.n_slots  # solution length (u16 elements)
        self.elite = torch.zeros((self.len + 1) & ~1, dtype=torch.int16, device=dev)  # even: reduced as int32
        ptr0, stride = eng.chain_device_ptr(0)
        self.rows_all = device_view(ptr0, (chains_per_rank, stride), "<i2", dev)
        self.kbuf = torch.zeros(1, dtype=torch.int64, device=dev)
        self.reduced_key = None   # device tensor after sync_device()
        self.best_score = None
        self.best_chain = None

    def sync_device(self):
        """min-allreduce + elite delivery entirely on the device (no host sync); read the result
        later with result().  Four stream-ordered operations: key copy, min-all-reduce, the
        library's owner-masked gather kernel (cs_*_exchange_select), sum-all-reduce of the elite."""
        self.kbuf.copy_(self.key)
        self.dist.all_reduce(self.kbuf, op=self.dist.ReduceOp.MIN)
        self.eng.exchange_select(self.kbuf.data_ptr(), self.elite.data_ptr(), self.elite.numel())
        self.dist.all_reduce(self.elite.view(torch.int32), op=self.dist.ReduceOp.SUM)
        self.reduced_key = self.kbuf
        return self.reduced_key

    def sync_device_torch_ops(self):
        """The same exchange written with torch ops only (what the gloo CPU tests run)."""
        self.reduced_key = exchange_best_device(self.dist, self.key, self.rows_all, self.elite, self.rank,
                                                self.cpr)
        return self.reduced_key

    def result(self):
        """(score part, global chain id) of the last sync_device(); this is the only host read."""
        k = int(self.reduced_key.item())
        self.best_score, self.best_chain = k >> 32, k & 0xFFFFFFFF
        return self.best_score, self.best_chain

    def _rows_of(self, local_chain: int) -> torch.Tensor:
        ptr, _ = self.eng.chain_device_ptr(local_chain)
        return device_view(ptr, (self.len,), "<i2", self.dev)

    def sync(self, inject_into_worst: bool = False):
        self.best_score, self.best_chain = exchange_best(self.dist, self.key, self._rows_of,
                                                         self.elite[: self.len], self.rank, self.cpr)
        if inject_into_worst and self.best_chain // self.cpr != self.rank and hasattr(self.eng, "set_chain_from_device"):
            worst = int(self.eng.scores().argmax())
            self.eng.set_chain_from_device(worst, self.elite.data_ptr())
        return self.best_score, self.best_chain


class PartitionedBoard:
    """ONE very large n-queens instance whose swap neighbourhood is split across ranks
    (SURVEY 8e, config 5).  Every rank holds a replica (same seed / same rows) and scans a
    triangular-balanced column slice; per step the 8-byte packed keys are min-all-reduced and
    every replica applies the same move, so no state ever crosses NVLink."""

    def __init__(self, eng, dist, rank: int, world: int):
        self.eng, self.dist, self.rank, self.world = eng, dist, rank, world
        _bind_to_torch_stream(eng)
        eng.set_partition(rank, world)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.key = device_view(eng.part_key_device_ptr(), (1,), "<i8", dev)

    def step(self):
        self.eng.part_scan()                       # enumerate + delta-score this slice
        if self.world > 1:
            self.dist.all_reduce(self.key, op=self.dist.ReduceOp.MIN)  # in place, on device
        return self.eng.part_apply()               # every replica accepts the winner
