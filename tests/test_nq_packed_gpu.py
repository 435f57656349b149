"""GPU parity suite for the packed-window fast scan (nq_step_kernel_v2): every delta it produces
equals the oracle's clone + full re-score, and its trajectories equal the scalar path's."""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _long_diagonal_board(n, k, rng):
    """permutation whose main diagonal holds exactly k queens (line count k), rest shuffled"""
    rows = np.arange(n)
    tail = rows[k:].copy()
    while True:
        rng.shuffle(tail)
        if not np.any(tail == np.arange(k, n)):  # no extra queen on the main diagonal
            break
    rows[k:] = tail
    return rows.astype(np.int64)


def test_every_delta_of_the_packed_scan_equals_full_rescore_difference():
    rng = np.random.default_rng(31)
    for n in (256, 257, 300, 384, 391):
        boards = [rng.permutation(n).astype(np.int64), _long_diagonal_board(n, 62, rng),
                  _long_diagonal_board(n, 63, rng)]  # 62: still packed; 63: scalar fallback
        with cs.NQueensChains(n, len(boards)) as e:
            e.set_chains(np.stack(boards))
            for k, rows in enumerate(boards):
                dev = e.neighbourhood_deltas(k)
                ref = orc.nq_neighbourhood_deltas(rows, orc.SWAP)
                bad = np.nonzero(dev != ref)[0]
                assert bad.size == 0, (n, k, bad[:5], dev[bad[:5]], ref[bad[:5]])


def test_packed_and_scalar_paths_walk_the_same_trajectory():
    for n, steps in [(256, 30), (1000, 12), (2049, 8), (10_000, 3), (12_096, 2)]:
        start = orc.nq_init_perm(5, 1, n)
        with cs.NQueensChains(n, 2, trace_capacity=32) as a, \
                cs.NQueensChains(n, 2, trace_capacity=32, force_scalar=True) as b:
            a.set_chains(np.stack([start, start[::-1].copy()]))
            b.set_chains(np.stack([start, start[::-1].copy()]))
            sa, sb = a.step(steps), b.step(steps)
            for k in range(2):
                ma, ca, ta = a.trace(k)
                mb, cb, tb = b.trace(k)
                assert ta == tb and np.array_equal(ma, mb) and np.array_equal(ca, cb), (n, k)
            assert np.array_equal(a.get_chains(), b.get_chains())
            assert sa.moves_scored == sb.moves_scored == 2 * steps * n * (n - 1) // 2
            assert a.score_full(0) == int(a.scores()[0]) == int(b.scores()[0])


def test_packed_path_local_search_and_replay_against_the_oracle():
    n = 300
    start = orc.nq_init_perm(9, 0, n)
    ref = orc.nq_local_search(start, allow_no_improvement_for=3, max_iterations=6, trace_cap=16)
    with cs.NQueensChains(n, 1, trace_capacity=16) as e:
        e.set_chains(start)
        e.local_search(3, 6)
        mv, sc, total = e.trace(0)
        assert total == ref["steps"] and np.array_equal(sc, ref["trace_score"])
        assert np.array_equal(mv[:, 0], ref["trace_a"]) and np.array_equal(mv[:, 1], ref["trace_b"])
        best, bsc = e.get_best_chains()
        assert int(bsc[0]) == ref["best_score"] and np.array_equal(best[0], ref["best"])


def test_non_permutation_chain_uses_the_scalar_fallback_inside_v2():
    n = 320
    rng = np.random.default_rng(2)
    rows = rng.permutation(n).astype(np.int64)
    rows[10] = rows[200]
    with cs.NQueensChains(n, 1, trace_capacity=4) as e:
        e.set_chains(rows)
        dev = e.neighbourhood_deltas(0)
        ref = orc.nq_neighbourhood_deltas(rows, orc.SWAP)
        assert np.array_equal(dev, ref)
        e.step(2)
        mv, sc, _ = e.trace(0)
        r = rows.copy()
        for (i, j), s in zip(mv, sc):
            r[i], r[j] = r[j], r[i]
            assert orc.nq_score(r) == int(s)
