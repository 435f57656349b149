"""GPU parity suite for the n-queens path: the CUDA kernels (through the C ABI) against the
CPU oracle's clone + full re-score (the reference's own formulation).  Bit-exact: scores are
integers (examples/nqueens/src/lib.rs:13,64)."""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

SIZES = [2, 3, 4, 5, 8, 31, 32, 33, 63, 64, 65, 97, 130]


def _states(n, rng):
    """permutation, random multiset (non-permutation), and degenerate boards"""
    out = [rng.permutation(n), rng.integers(0, n, size=n), np.arange(n), np.arange(n)[::-1].copy(),
           np.zeros(n, dtype=np.int64)]
    if n >= 4:
        r = rng.permutation(n)
        r[1] = r[0]  # a single duplicate row
        out.append(r)
    return [np.ascontiguousarray(x, dtype=np.int64) for x in out]


def test_init_matches_host_philox_mirror_and_scores():
    for n in (1, 2, 17, 64, 300):
        with cs.NQueensChains(n, 5, seed=99, chain_offset=11) as e:
            e.init_random()
            rows, sc = e.get_chains(), e.scores()
            for k in range(5):
                assert np.array_equal(rows[k], orc.nq_init_perm(99, 11 + k, n))
                assert int(sc[k]) == orc.nq_score(rows[k])
                assert e.score_full(k) == int(sc[k])


def test_golden_vectors_on_device(golden_dir):
    import json, os
    g = json.load(open(os.path.join(golden_dir, "nq_kat.json")))
    for case in g["reference"] + g["derived"]:
        rows = np.array(case["rows"], dtype=np.int64)
        with cs.NQueensChains(len(rows), 1) as e:
            e.set_chains(rows)
            assert e.score_full(0) == case["score"]  # pair-loop kernel
            assert int(e.scores()[0]) == case["score"]  # counter formulation
            assert np.array_equal(e.get_chains()[0], rows)


@pytest.mark.parametrize("kind", [cs.SWAP, cs.CHANGE])
def test_every_delta_equals_full_rescore_difference(kind):
    """Every candidate of the PRODUCTION scan vs. the oracle's clone + full re-score."""
    rng = np.random.default_rng(2024 + kind)
    for n in SIZES:
        states = _states(n, rng)
        with cs.NQueensChains(n, len(states), neighbourhood=kind) as e:
            e.set_chains(np.stack(states))
            for k, rows in enumerate(states):
                dev = e.neighbourhood_deltas(k)
                ref = orc.nq_neighbourhood_deltas(rows, kind)
                assert dev.shape == ref.shape
                bad = np.nonzero(dev != ref)[0]
                assert bad.size == 0, (n, k, bad[:5], dev[bad[:5]], ref[bad[:5]])
                assert int(e.scores()[k]) == orc.nq_score(rows)
                assert e.score_full(k) == orc.nq_score(rows)


@pytest.mark.parametrize("kind", [cs.SWAP, cs.CHANGE])
def test_eval_moves_hook_and_enumerate(kind):
    rng = np.random.default_rng(7)
    for n in (6, 40):
        for rows in _states(n, rng)[:3]:
            with cs.NQueensChains(n, 1, neighbourhood=kind) as e:
                e.set_chains(rows)
                mv = e.enumerate(0)
                full = orc.nq_neighbourhood_deltas(rows, kind)
                assert len(mv) == int((full != orc.INT64_MAX).sum())
                d = e.eval_moves(mv[:, 0], mv[:, 1], 0, kind)
                assert np.array_equal(d, orc.nq_eval_moves(rows, mv[:, 0], mv[:, 1], kind))
                # identity moves report INT64_MAX on both sides
                if kind == cs.CHANGE:
                    a, b = np.arange(n), rows
                    assert (e.eval_moves(a, b, 0, kind) == orc.INT64_MAX).all()


def test_eval_moves_large_board_sample():
    n = 2000
    rng = np.random.default_rng(3)
    rows = rng.permutation(n).astype(np.int64)
    rows[5] = rows[900]  # non-permutation
    a = rng.integers(0, n, size=300)
    b = rng.integers(0, n, size=300)
    keep = a != b
    a, b = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
    with cs.NQueensChains(n, 1) as e:
        e.set_chains(rows)
        assert np.array_equal(e.eval_moves(a, b, 0, cs.SWAP), orc.nq_eval_moves(rows, a, b, orc.SWAP))
        v = rng.integers(0, n, size=len(a))
        assert np.array_equal(e.eval_moves(a, v, 0, cs.CHANGE), orc.nq_eval_moves(rows, a, v, orc.CHANGE))


@pytest.mark.parametrize("kind", [cs.SWAP, cs.CHANGE])
def test_step_picks_the_oracle_argmin_and_trace_replays(kind):
    """Chosen move == first minimum of the oracle's full re-score in (a, b) order; replaying
    the device's trace through the oracle scorer reproduces the score trajectory."""
    rng = np.random.default_rng(11)
    for n in (5, 8, 24, 33, 70):
        starts = np.stack(_states(n, rng)[:4])
        with cs.NQueensChains(n, len(starts), neighbourhood=kind, trace_capacity=16) as e:
            e.set_chains(starts)
            e.step(6)
            after = e.get_chains()
            scores = e.scores()
            for k in range(len(starts)):
                mv, sc, total = e.trace(k)
                assert total == len(mv) <= 6
                r = starts[k].copy()
                for (a, b), s in zip(mv, sc):
                    if orc.nq_score(r) == 0:
                        pytest.fail("stepped past is_best")
                    d = orc.nq_neighbourhood_deltas(r, kind)
                    best = int(np.argmin(d))  # first minimum in enumeration order
                    if kind == cs.SWAP:
                        ii, jj = np.triu_indices(n, 1)
                        assert (int(a), int(b)) == (int(ii[best]), int(jj[best])), (n, k)
                        r[a], r[b] = r[b], r[a]
                    else:
                        assert (int(a), int(b)) == (best // n, best % n), (n, k)
                        r[a] = b
                    assert orc.nq_score(r) == int(s)
                assert np.array_equal(r, after[k])
                assert orc.nq_score(r) == int(scores[k])


@pytest.mark.parametrize("kind", [cs.SWAP, cs.CHANGE])
def test_local_search_matches_oracle_execute(kind):
    """LocalSearch::execute (local_search.rs:301-342): same trajectory, same returned best."""
    for n, seed, allow, iters in [(8, 1, 5, 100), (12, 2, 3, 100), (20, 3, 5, 7), (30, 4, 2, 1000), (16, 5, 1, 50)]:
        start = orc.nq_init_perm(seed, 0, n)
        ref = orc.nq_local_search(start, kind=kind, tie=orc.TIE_MOVE_ORDER,
                                  allow_no_improvement_for=allow, max_iterations=iters, trace_cap=2048)
        with cs.NQueensChains(n, 3, neighbourhood=kind, trace_capacity=2048) as e:
            e.set_chains(np.stack([start, start[::-1].copy(), start]))
            st = e.local_search(allow, iters)
            mv, sc, total = e.trace(0)
            assert total == ref["steps"]
            assert np.array_equal(mv[:, 0], ref["trace_a"]) and np.array_equal(mv[:, 1], ref["trace_b"])
            assert np.array_equal(sc, ref["trace_score"])
            best, best_sc = e.get_best_chains()
            assert int(best_sc[0]) == ref["best_score"] and np.array_equal(best[0], ref["best"])
            assert np.array_equal(e.get_chains()[0], ref["current"])
            assert np.array_equal(best[2], best[0])  # identical chains are deterministic
            status = e.status()
            if ref["best_score"] == 0 and ref["current_score"] == 0:
                assert status[0] == L.CHAIN_BEST
            assert st.steps_accepted >= ref["steps"]
        # the single-solution, execute()-shaped entry point
        ls = cs.LocalSearch(cs.NQueensMoveProposer(n, kind), cs.NQueensSolutionScoreCalculator(), iters)
        out = ls.execute(cs.NQueensSolution(start), allow)
        assert out.score.value == ref["best_score"] and np.array_equal(out.solution.rows, ref["best"])


def test_edge_cases_and_errors():
    with cs.NQueensChains(1, 2) as e:  # n = 1: empty neighbourhood, score 0
        e.set_chains(np.zeros((2, 1), dtype=np.int64))
        st = e.step(3)
        assert st.moves_scored == 0 and st.best_score == 0
    with cs.NQueensChains(4, 1) as e:
        with pytest.raises(cs.CsError):  # stepping before any solution exists
            e.step(1)
        e.set_chains([0, 0, 0, 0])  # every swap is an identity move -> empty neighbourhood
        st = e.step(2)
        assert st.steps_accepted == 0 and e.status()[0] == L.CHAIN_EMPTY and int(e.scores()[0]) == 12
        e.set_chains([1, 3, 0, 2])  # already is_best
        st = e.step(2)
        assert st.steps_accepted == 0 and e.status()[0] == L.CHAIN_BEST
        with pytest.raises(cs.CsError) as err:
            e.set_chains([0, 1, 2, 4])  # row outside the board
        assert err.value.status == L.CS_ERR_INVALID_ARG
        with pytest.raises(cs.CsError):
            e.get_chains(1, 1)  # chain out of range
        with pytest.raises(cs.CsError):
            e.eval_moves([0], [9], 0, cs.SWAP)
    lib = cs.load()
    import ctypes as C
    h = C.c_void_p()
    cfg = L.CsNqConfig(n=8, n_chains=1, chain_offset=0, trace_capacity=0, seed=1, device=999, neighbourhood=0, flags=0)
    assert lib.cs_nq_create(C.byref(cfg), C.byref(h)) == L.CS_ERR_INVALID_ARG


def test_moves_scored_accounting():
    n, chains = 50, 7
    with cs.NQueensChains(n, chains, seed=5) as e:
        e.init_random()
        st = e.step(1)
        assert st.moves_scored == chains * n * (n - 1) // 2
        assert st.steps_accepted == chains
        assert st.kernel_launches >= 1 and st.device_ms > 0
    with cs.NQueensChains(6, 1) as e:
        e.set_chains([0, 0, 1, 2, 3, 3])  # two identity pairs
        assert e.step(1).moves_scored == 15 - 2


def test_full_size_board_properties():
    """BASELINE config 2 board size (n = 10 000): size-independent properties -- delta-tracked
    score == device pair-loop re-score == oracle re-score after replaying the trace."""
    n, chains = 10_000, 3
    with cs.NQueensChains(n, chains, seed=42, trace_capacity=8) as e:
        e.init_random()
        rows0 = e.get_chains()
        assert np.array_equal(rows0[1], orc.nq_init_perm(42, 1, n))
        s0 = e.scores()
        assert int(s0[0]) == orc.nq_score(rows0[0])
        st = e.step(2)
        assert st.moves_scored == 2 * chains * n * (n - 1) // 2
        rows1, s1 = e.get_chains(), e.scores()
        for k in range(chains):
            mv, sc, total = e.trace(k)
            assert total == 2
            r = rows0[k].copy()
            for (i, j), s in zip(mv, sc):
                assert i < j
                r[i], r[j] = r[j], r[i]
            assert np.array_equal(r, rows1[k])
            assert sorted(r.tolist()) == list(range(n))  # swaps keep the permutation
            assert e.score_full(k) == int(s1[k]) == int(sc[-1])
            assert int(sc[-1]) < int(s0[k])  # a random board always has an improving swap
        assert orc.nq_score(rows1[0]) == int(s1[0])
        # sampled deltas at full size against clone + full re-score
        rng = np.random.default_rng(0)
        a = rng.integers(0, n, 24)
        b = rng.integers(0, n, 24)
        keep = a != b
        a, b = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
        assert np.array_equal(e.eval_moves(a, b, 0, cs.SWAP), orc.nq_eval_moves(rows1[0], a, b, orc.SWAP))


def test_max_smem_board_runs():
    n = L.CS_NQ_MAX_N_SMEM
    with cs.NQueensChains(n, 1, seed=1, trace_capacity=2) as e:
        e.init_random()
        s0 = int(e.scores()[0])
        assert e.score_full(0) == s0
        e.step(1)
        assert e.score_full(0) == int(e.scores()[0]) < s0


def test_async_staging_equals_synchronous_set_chains():
    """cs_nq_set_chains_async + cs_nq_commit_chains: a pipelined stream of batches gives exactly
    what cs_nq_set_chains gives, the copy may overlap a running step, and misuse is refused."""
    import torch

    n, chains = 300, 24
    rng = np.random.default_rng(3)
    batches = [np.stack([rng.permutation(n) for _ in range(chains)]).astype(np.int64) for _ in range(3)]
    pinned = [torch.from_numpy(b).pin_memory() for b in batches]
    with cs.NQueensChains(n, chains, trace_capacity=4) as a, cs.NQueensChains(n, chains, trace_capacity=4) as b:
        with pytest.raises(cs.CsError):
            a.commit_chains()                               # nothing pending
        a.set_chains_async_ptr(pinned[0].data_ptr(), chains)
        with pytest.raises(cs.CsError):
            a.set_chains_async_ptr(pinned[1].data_ptr(), chains)   # one pending upload per handle
        for k in range(3):
            a.commit_chains()
            if k + 1 < 3:
                a.set_chains_async_ptr(pinned[k + 1].data_ptr(), chains)   # overlaps the step below
            sa = a.step(3)
            b.set_chains(batches[k])
            sb = b.step(3)
            assert np.array_equal(a.get_chains(), b.get_chains()) and np.array_equal(a.scores(), b.scores())
            assert sa.moves_scored == sb.moves_scored and sa.best_score == sb.best_score
            for c in (0, chains - 1):
                ma, ca, ta = a.trace(c)
                mb, cb_, tb = b.trace(c)
                assert ta == tb and np.array_equal(ma, mb) and np.array_equal(ca, cb_)
        bad = torch.full((chains, n), n, dtype=torch.int64).pin_memory()   # row value out of range
        a.set_chains_async_ptr(bad.data_ptr(), chains)
        with pytest.raises(cs.CsError):
            a.commit_chains()
    with cs.NQueensChains(64, 1, force_global=True) as g:
        with pytest.raises(cs.CsError):
            g.set_chains_async_ptr(pinned[0].data_ptr(), 1)   # big-board path: unsupported
