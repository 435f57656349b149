import sys, os, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import constraint_solver_b200 as cs
n = 1_000_000
eng = cs.NQueensChains(n, 1, seed=42)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.init_random()
for parts in (8,):
    for k in range(parts):
        eng.set_partition(k, parts)
        eng.part_scan(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.part_scan(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(parts, k, "%.2f ms" % (dt * 1e3))
