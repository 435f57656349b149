import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import constraint_solver_b200 as cs
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import es_instance
for name in ("es2000", "es50"):
    w, ids, hol = es_instance(name, 42)
    D, chains = w["D"], w["chains"]
    eng = cs.ScheduleChains(D, ids, holidays=hol, n_chains=chains, seed=42)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.init_random()
    host = torch.empty((chains, D + 1), dtype=torch.int64, pin_memory=True)
    host.copy_(torch.from_numpy(np.ascontiguousarray(eng.get_chains())))
    def T(f, reps=20):
        f(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): f()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3 / reps
    print(name, "set_chains %.3f ms  step %.3f ms  scores %.3f ms" % (
        T(lambda: eng.set_chains_ptr(host.data_ptr(), chains)), T(lambda: eng.step(1)), T(lambda: eng.scores())))
