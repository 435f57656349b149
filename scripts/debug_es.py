import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import constraint_solver_b200 as cs
from oracle import oracle as orc
sys.path.insert(0, "tests")
from test_es_gpu import _oracle_deltas
rng = np.random.default_rng(4)
D, ids, wd, hol = 31, np.arange(7), 0, [(0, 0), (0, 1), (3, 3)]
a = ids[rng.integers(0, 7, size=D + 1)]
print("a", a.tolist())
with cs.ScheduleChains(D, ids, start_weekday=wd, holidays=hol) as e:
    e.set_chains(a)
    mv = e.enumerate(0)
    ref_h, ref_s = _oracle_deltas(a[:D], ids, wd, hol)
    dev_h, dev_s = e.neighbourhood_deltas(0)
    bad = np.nonzero((dev_h != ref_h) | (dev_s != ref_s))[0]
    print("scan bad", bad[:20], dev_s[bad[:20]], ref_s[bad[:20]])
    dh, ds = e.eval_moves(mv["kind"], mv["a"], mv["b"], 0)
    keep = ref_h != orc.INT64_MAX
    bad = np.nonzero((dh != ref_h[keep]) | (ds != ref_s[keep]))[0]
    print("eval bad", len(bad), bad[:20])
    for b in bad[:12]:
        print(mv[b], "got", dh[b], ds[b], "ref", ref_h[keep][b], ref_s[keep][b])
