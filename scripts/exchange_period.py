"""Cost of the best-score exchange at the scheduling step: 56 days x 2000
employees x 4096 chains per GPU, the exchange run every 1, 8 and 64 chain-steps: host-sync
(key.item() + broadcast), device-side (BestExchange.sync_device: min-all-reduce, the library's
select kernel, sum-all-reduce; no host read) after a blocking cs_es_step, and device-side after
cs_es_step_enqueue (no host synchronisation anywhere in the loop).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/exchange_period.py > gpurun_out/exchange_period.json

Device-timed (CUDA events on the stream everything runs on), max over ranks; rank 0 prints one
JSON line.  Not a bench value: it explains the N > 1 lines of bench.py.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch
import torch.distributed as dist

import bench
import constraint_solver_b200 as cs
from constraint_solver_b200.dist import BestExchange

TOTAL_STEPS = 256


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, ids, hol, _ = bench.es_instance("es2000", 42)
    chains = w["chains"]
    eng = cs.ScheduleChains(w["D"], ids, holidays=hol, n_chains=chains, seed=42, chain_offset=rank * chains,
                            device=local)
    eng.init_random()
    xchg = BestExchange(eng, dist, rank, world, chains)   # binds the engine to torch's current stream
    start = eng.get_chains()
    stream = torch.cuda.current_stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(period, mode):
        eng.set_chains(start)
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(TOTAL_STEPS // period):
            if mode.endswith("_async"):
                eng.step_enqueue(period)      # the host never waits inside the loop
            else:
                eng.step(period)              # cs_es_step: waits for the launch and reads its stats
            if mode.startswith("device"):
                xchg.sync_device()
            elif mode == "host":
                xchg.sync()
        ev1.record(stream)
        if mode.endswith("_async"):
            eng.step_wait()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / TOTAL_STEPS

    modes = ("none", "none_async", "host", "device", "device_async")
    rows = []
    for mode in modes:
        timed(64, mode)   # warm-up (NCCL channels, first launches)
    for period in (1, 8, 64):
        t = {m: timed(period, m) for m in modes}
        row = {"exchange_every_steps": period}
        row.update({f"ms_per_chain_step_{m}": v for m, v in t.items()})
        row["exchange_cost_ms"] = {"host_sync": (t["host"] - t["none"]) * period,
                                   "device": (t["device"] - t["none"]) * period,
                                   "device_async": (t["device_async"] - t["none_async"]) * period}
        rows.append(row)
    if rank == 0:
        print(json.dumps({"workload": "scheduling 56 x 2000, 4096 chains per GPU", "n_gpus": world,
                          "chain_steps_timed": TOTAL_STEPS, "rows": rows}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
