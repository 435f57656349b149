"""Small run that launches every kernel once (for compute-sanitizer memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import constraint_solver_b200 as cs  # noqa: E402

rng = np.random.default_rng(0)
# scalar shared-memory path, both neighbourhoods, perm + non-perm
for kind in (cs.SWAP, cs.CHANGE):
    with cs.NQueensChains(70, 6, neighbourhood=kind, trace_capacity=8) as e:
        e.init_random()
        rows = e.get_chains()
        rows[1, 3] = rows[1, 9]
        e.set_chains(rows)
        e.neighbourhood_deltas(0); e.neighbourhood_deltas(1)
        e.step(3); e.local_search(3, 5); e.score_full(0)
        e.eval_moves([0, 1], [5, 9], 0, kind); e.enumerate(0); e.trace(0); e.best()
        e.ils_init(4, 8); e.ils_run(3, 6, 3); e.ils_best(0); e.ils_log(0)
# packed path (n >= 256) incl. the head/tail chunk masks, and its scalar fallback
for n in (256, 391, 1100):
    with cs.NQueensChains(n, 3, trace_capacity=4) as e:
        e.init_random()
        rows = e.get_chains()
        rows[2, 1] = rows[2, 0]
        e.set_chains(rows)
        if n < 400:
            e.neighbourhood_deltas(0); e.neighbourhood_deltas(2)
        e.step(2)
with cs.NQueensChains(300, 2, force_scalar=True) as e:
    e.init_random(); e.step(2)
# big-board path + partitions
with cs.NQueensChains(200, 1, force_global=True, trace_capacity=4) as e:
    e.init_random(); e.neighbourhood_deltas(0); e.step(2); e.local_search(2, 3)
    e.set_partition(1, 3); e.part_scan(); e.part_apply(); e.eval_moves([0], [7], 0, cs.SWAP)
with cs.NQueensChains(17000, 1) as e:
    e.init_random(); e.step(1); e.score_full(0)
# scheduling
for D, E in ((9, 3), (31, 7), (64, 70)):
    ids = np.arange(E) * 2
    hol = [(int(ids[0]), 1), (int(ids[-1]), D - 1)]
    with cs.ScheduleChains(D, ids, start_weekday=3, holidays=hol, n_chains=5, trace_capacity=4) as e:
        e.init_random(); e.neighbourhood_deltas(0); e.step(3); e.local_search(3, 4)
        e.score_full(0); e.enumerate(0); e.eval_moves([0, 1], [1, 2], [0, 3], 0); e.trace(0); e.best()
        e.ils_init(3, 4); e.ils_run(3, 5, 3); e.ils_best(0)
print("sanitize smoke done")
