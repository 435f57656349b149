import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import constraint_solver_b200 as cs
n, chains = 10_000, 4096
eng = cs.NQueensChains(n, chains, seed=42)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.init_random()
host = torch.empty((chains, n), dtype=torch.int64, pin_memory=True)
host.copy_(torch.from_numpy(eng.get_chains()))
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
print("sync set_chains %.2f ms" % T(lambda: eng.set_chains_ptr(host.data_ptr(), chains))[0])
print("step            %.2f ms" % T(lambda: eng.step(1))[0])
print("scores          %.2f ms" % T(lambda: eng.scores())[0])
print("async enqueue   %.2f ms (includes the copy: T() synchronises the device)" % T(lambda: eng.set_chains_async_ptr(host.data_ptr(), chains))[0])
print("commit          %.2f ms" % T(lambda: eng.commit_chains())[0])
# overlapped: enqueue then step, no sync in between
torch.cuda.synchronize(); t0 = time.perf_counter()
eng.set_chains_async_ptr(host.data_ptr(), chains); t1 = time.perf_counter()
st = eng.step(1); t2 = time.perf_counter()
eng.commit_chains(); t3 = time.perf_counter()
print("overlapped: enqueue %.2f ms, step %.2f ms (device %.2f), commit %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st.device_ms, (t3 - t2) * 1e3))
# the bench's e2e loop, iteration by iteration
torch.cuda.synchronize(); t0 = time.perf_counter()
eng.set_chains_async_ptr(host.data_ptr(), chains)
for k in range(5):
    a = time.perf_counter(); eng.commit_chains()
    b = time.perf_counter()
    if k + 1 < 5: eng.set_chains_async_ptr(host.data_ptr(), chains)
    c = time.perf_counter(); st = eng.step(1)
    d = time.perf_counter(); sc = eng.scores()
    e = time.perf_counter()
    print("iter %d: commit %.2f  enqueue %.2f  step %.2f (device %.2f)  scores %.2f ms" % (k, (b-a)*1e3, (c-b)*1e3, (d-c)*1e3, st.device_ms, (e-d)*1e3))
torch.cuda.synchronize(); print("total %.2f ms" % ((time.perf_counter() - t0) * 1e3))
