"""GPU parity suite for the ILS shell: the device chains replay bit-identically through the
oracle's restatement of IteratedLocalSearch::execute_round (iterated_local_search.rs:173-202) --
same per-round local minimum, same acceptance choice, same best."""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", [cs.CHANGE, cs.SWAP])
def test_nq_ils_replays_through_the_oracle(kind):
    # rounds > 50 exercises the random restart; cap 3 exercises best-set eviction + dedup
    for n, cap, rounds, iters, allow in [(8, 32, 120, 40, 5), (12, 3, 70, 25, 5), (20, 2, 55, 12, 3)]:
        chains = 5
        with cs.NQueensChains(n, chains, seed=42, chain_offset=3, neighbourhood=kind) as e:
            e.init_random()
            e.ils_init(cap, log_capacity=rounds)
            st = e.ils_run(rounds, iters, allow)
            for k in range(chains):
                ref = orc.nq_ils(42, 3 + k, n, kind=kind, ls_max_iterations=iters,
                                 allow_no_improvement_for=allow, rounds=rounds, best_cap=cap)
                key, choice, total = e.ils_log(k)
                assert total == ref["rounds"], (n, k, total, ref["rounds"])
                assert np.array_equal(key, ref["round_new_score"]), (n, k)
                assert np.array_equal(choice, ref["round_choice"]), (n, k)
                rows, sc = e.ils_best(k)
                assert sc == ref["best_score"] and np.array_equal(rows, ref["best"])
                assert orc.nq_score(rows) == sc
                assert np.array_equal(e.get_chains(k, 1)[0], ref["current"])
            assert st["rounds_run"] == rounds
            bests = [e.ils_best(k)[1] for k in range(chains)]
            assert st["best_key"] == min(bests) and st["chains_done"] == sum(b == 0 for b in bests)


def test_nq_ils_reference_repeatable_property():
    """examples/nqueens/src/main.rs:157-200: n = 8 reaches score 0 (here: every chain, with the
    reference's change moves and its solver constants) and identical runs give identical results."""
    outs = []
    for _ in range(2):
        with cs.NQueensChains(8, 16, seed=42, neighbourhood=cs.CHANGE) as e:
            e.init_random()
            e.ils_init(32)
            st = e.ils_run(400, 10_000, 5)
            outs.append([e.ils_best(k) for k in range(16)])
            assert st["chains_done"] == 16
    for (r1, s1), (r2, s2) in zip(*outs):
        assert s1 == s2 == 0 and np.array_equal(r1, r2) and orc.nq_score(r1) == 0


def test_nq_ils_stop_when_any_best_and_errors():
    with cs.NQueensChains(10, 64, seed=7, neighbourhood=cs.CHANGE) as e:
        with pytest.raises(cs.CsError):
            e.ils_init(32)  # no solution yet
        e.init_random()
        with pytest.raises(cs.CsError):
            e.ils_run(1, 10, 5)  # ils_init not called
        e.ils_init(8)
        with pytest.raises(cs.CsError) as err:
            e.ils_best(0)  # get_best before any round: the reference unwrap()s None
        assert err.value.status == L.CS_ERR_STATE
        st = e.ils_run(500, 1000, 5, stop_when_any_best=True)
        assert st["chains_done"] >= 1 and st["best_key"] == 0 and st["rounds_run"] < 500
        rows, sc = e.ils_best(st["best_chain"])
        assert sc == 0 and orc.nq_score(rows) == 0
        with pytest.raises(cs.CsError):
            e.ils_init(65)


def test_es_ils_replays_through_the_oracle():
    for D, E, cap, rounds, iters, allow in [(14, 4, 64, 60, 15, 20), (10, 3, 2, 55, 10, 4), (21, 5, 4, 12, 8, 20)]:
        ids = np.arange(E) * 2 + 1
        hol = [(int(ids[0]), 1), (int(ids[E - 1]), D - 2)]
        chains = 3
        with cs.ScheduleChains(D, ids, start_weekday=2, holidays=hol, n_chains=chains, seed=11, chain_offset=5) as e:
            e.init_random()
            e.ils_init(cap, log_capacity=rounds)
            e.ils_run(rounds, iters, allow)
            for k in range(chains):
                ref = orc.es_ils(11, 5 + k, D, ids, 2, hol, ls_max_iterations=iters,
                                 allow_no_improvement_for=allow, rounds=rounds, best_cap=cap)
                key, choice, total = e.ils_log(k)
                assert total == ref["rounds"]
                assert np.array_equal(key, ref["round_new_key"]), (D, k, key[:8], ref["round_new_key"][:8])
                assert np.array_equal(choice, ref["round_choice"])
                rows, hard, soft = e.ils_best(k)
                assert (hard, soft) == (ref["best_hard"], ref["best_soft"])
                assert np.array_equal(rows, ref["best"])
                assert orc.es_score(rows[:D], 2, hol) == (hard, soft)
