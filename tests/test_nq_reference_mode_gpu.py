"""GPU parity for REFERENCE MODE: the reference's own n-queens proposer
(examples/nqueens/src/lib.rs:177-255), window (local_search.rs:321) and derived-Ord tie-break
(:29-37,323) on the device, against the oracle's literal restatement (full clones, full
re-score, lexicographic vector compare) -- identical trajectories from the same Philox stream."""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_reference_mode_local_search_trajectory_equals_oracle():
    for n, allow, iters, window in [(8, 5, 10_000, 40), (20, 5, 200, 100), (64, 5, 300, 320),
                                    (64, 2, 50, 7), (41, 3, 100, 1000), (130, 5, 60, 650)]:
        chains = 4
        with cs.NQueensChains(n, chains, seed=42, chain_offset=2, neighbourhood=cs.CHANGE,
                              trace_capacity=512, reference_proposer=True) as e:
            e.set_window(window)
            e.init_random()
            starts = e.get_chains()
            st = e.local_search(allow, iters)
            best, bsc = e.get_best_chains()
            cur = e.get_chains()
            for k in range(chains):
                ref = orc.nq_local_search_ref(starts[k], 42, 2 + k, 0, allow, iters, window, trace_cap=512)
                mv, sc, total = e.trace(k)
                assert total == ref["steps"], (n, k, total, ref["steps"])
                m = min(total, 512)
                assert np.array_equal(mv[:m, 0], ref["trace_a"][:m]) and np.array_equal(mv[:m, 1], ref["trace_b"][:m]), (n, k)
                assert np.array_equal(sc[:m], ref["trace_score"][:m])
                assert int(bsc[k]) == ref["best_score"] and np.array_equal(best[k], ref["best"])
                assert np.array_equal(cur[k], ref["current"])
            assert st.moves_scored > 0


def test_reference_mode_ils_replays_and_solves_like_the_reference_test():
    """examples/nqueens/src/main.rs:157-200 (`repeatable`): n = 8, its solver constants, seeds
    -> score 0 and run-to-run identical; here additionally identical to the CPU restatement."""
    n, chains = 8, 8
    with cs.NQueensChains(n, chains, seed=1234, neighbourhood=cs.CHANGE, reference_proposer=True) as e:
        e.set_window(5 * n)
        e.init_random()
        e.ils_init(32, log_capacity=256)
        st = e.ils_run(256, 10_000, 5)
        assert st["chains_done"] == chains
        for k in range(chains):
            ref = orc.nq_ils(1234, k, n, kind=orc.CHANGE, ls_max_iterations=10_000,
                             allow_no_improvement_for=5, rounds=256, best_cap=32, ref_window=5 * n)
            key, choice, total = e.ils_log(k)
            assert total == ref["rounds"]
            assert np.array_equal(key, ref["round_new_score"]) and np.array_equal(choice, ref["round_choice"])
            rows, sc = e.ils_best(k)
            assert sc == ref["best_score"] == 0 and np.array_equal(rows, ref["best"])


def test_reference_mode_config1_board():
    """BASELINE configs[0]: n = 64 with the reference's constants; a few chains, full replay."""
    n, chains = 64, 3
    with cs.NQueensChains(n, chains, seed=42, neighbourhood=cs.CHANGE, reference_proposer=True) as e:
        e.init_random()  # window defaults to 5 * n
        e.ils_init(32, log_capacity=64)
        e.ils_run(40, 10_000, 5)
        for k in range(chains):
            ref = orc.nq_ils(42, k, n, kind=orc.CHANGE, ls_max_iterations=10_000,
                             allow_no_improvement_for=5, rounds=40, best_cap=32, ref_window=5 * n)
            key, choice, total = e.ils_log(k)
            assert total == ref["rounds"] and np.array_equal(key, ref["round_new_score"])
            assert np.array_equal(choice, ref["round_choice"])
            rows, sc = e.ils_best(k)
            assert sc == ref["best_score"] and np.array_equal(rows, ref["best"])


def test_reference_mode_is_change_only():
    with pytest.raises(cs.CsError) as err:
        cs.NQueensChains(16, 1, neighbourhood=cs.SWAP, reference_proposer=True)
    assert err.value.status == L.CS_ERR_UNSUPPORTED
