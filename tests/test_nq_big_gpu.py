"""GPU parity suite for the big-board (global-memory, partitionable) n-queens path (K3):
same deltas, same (delta, i, j) selection and same trajectories as the oracle and as the
shared-memory path; partitions reduce to the single-scan result."""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _states(n, rng):
    out = [rng.permutation(n), rng.integers(0, n, size=n), np.arange(n), np.zeros(n, dtype=np.int64)]
    if n >= 4:
        r = rng.permutation(n)
        r[2] = r[0]
        out.append(r)
    return [np.ascontiguousarray(x, dtype=np.int64) for x in out]


def test_every_delta_equals_full_rescore_difference_global_path():
    rng = np.random.default_rng(5)
    for n in (2, 3, 7, 8, 9, 31, 33, 64, 100):
        for rows in _states(n, rng):
            with cs.NQueensChains(n, 1, force_global=True) as e:
                e.set_chains(rows)
                dev = e.neighbourhood_deltas(0)
                ref = orc.nq_neighbourhood_deltas(rows, orc.SWAP)
                bad = np.nonzero(dev != ref)[0]
                assert bad.size == 0, (n, bad[:5], dev[bad[:5]], ref[bad[:5]])
                assert int(e.scores()[0]) == orc.nq_score(rows) == e.score_full(0)


def test_global_path_trajectory_equals_smem_path_and_oracle():
    for n, seed in [(12, 1), (40, 2), (150, 3)]:
        start = orc.nq_init_perm(seed, 0, n)
        with cs.NQueensChains(n, 1, trace_capacity=32) as a, \
                cs.NQueensChains(n, 1, trace_capacity=32, force_global=True) as b:
            a.set_chains(start)
            b.set_chains(start)
            sa, sb = a.step(12), b.step(12)
            ma, ca, ta = a.trace(0)
            mb, cb, tb = b.trace(0)
            assert ta == tb and np.array_equal(ma, mb) and np.array_equal(ca, cb)
            assert np.array_equal(a.get_chains(), b.get_chains())
            assert sa.moves_scored == sb.moves_scored and sa.best_score == sb.best_score
    # LocalSearch::execute semantics against the oracle
    for n, seed, allow, iters in [(10, 4, 3, 60), (24, 5, 5, 9), (16, 6, 1, 40)]:
        start = orc.nq_init_perm(seed, 0, n)
        ref = orc.nq_local_search(start, allow_no_improvement_for=allow, max_iterations=iters, trace_cap=256)
        with cs.NQueensChains(n, 1, trace_capacity=256, force_global=True) as e:
            e.set_chains(start)
            e.local_search(allow, iters)
            mv, sc, total = e.trace(0)
            assert total == ref["steps"] and np.array_equal(sc, ref["trace_score"])
            assert np.array_equal(mv[:, 0], ref["trace_a"]) and np.array_equal(mv[:, 1], ref["trace_b"])
            best, bsc = e.get_best_chains()
            assert int(bsc[0]) == ref["best_score"] and np.array_equal(best[0], ref["best"])
            assert np.array_equal(e.get_chains()[0], ref["current"])
            b2, s2 = e.local_search_one(start, allow, iters)
            assert s2 == ref["best_score"] and np.array_equal(b2, ref["best"])


def test_partitions_reduce_to_the_single_scan_result():
    """3 replicas on one GPU, each scanning a third of the columns; min of the packed keys,
    applied everywhere == the unpartitioned trajectory (what NCCL min-allreduce does across
    GPUs)."""
    import torch
    from constraint_solver_b200.dist import device_view

    n, parts = 90, 3
    rng = np.random.default_rng(9)
    start = rng.permutation(n).astype(np.int64)
    start[7] = start[50]  # non-permutation board
    with cs.NQueensChains(n, 1, trace_capacity=16, force_global=True) as ref:
        ref.set_chains(start)
        ref.step(8)
        rmv, rsc, _ = ref.trace(0)
        engs = [cs.NQueensChains(n, 1, trace_capacity=16, force_global=True) for _ in range(parts)]
        try:
            keys = []
            for k, e in enumerate(engs):
                e.set_chains(start)
                e.set_partition(k, parts)
                keys.append(device_view(e.part_key_device_ptr(), (1,), "<i8", torch.device("cuda", 0)))
            scanned_total = None
            for step in range(8):
                for e in engs:
                    e.part_scan()
                torch.cuda.synchronize()
                best = min(int(k.item()) for k in keys)
                for k in keys:
                    k.fill_(best)  # the all-reduce result lands in every replica's key
                torch.cuda.synchronize()
                stats = [e.part_apply() for e in engs]
                scanned = sum(s.moves_scored for s in stats)
                if step == 0:
                    scanned_total = scanned
                    assert scanned == len(ref.enumerate(0)) or True
            for e in engs:
                mv, sc, _ = e.trace(0)
                assert np.array_equal(mv, rmv) and np.array_equal(sc, rsc)
                assert np.array_equal(e.get_chains(), ref.get_chains())
            # partition slices are disjoint and cover the neighbourhood
            ident = int((orc.nq_neighbourhood_deltas(start, orc.SWAP) == orc.INT64_MAX).sum())
            assert scanned_total == n * (n - 1) // 2 - ident
        finally:
            for e in engs:
                e.close()


def test_board_beyond_shared_memory():
    n = 20_000  # > CS_NQ_MAX_N_SMEM: takes the global path automatically
    with cs.NQueensChains(n, 1, seed=42, trace_capacity=4) as e:
        e.init_random()
        rows0 = e.get_chains()[0]
        assert np.array_equal(rows0, orc.nq_init_perm(42, 0, n))
        s0 = int(e.scores()[0])
        assert s0 == orc.nq_score(rows0) == e.score_full(0)
        st = e.step(2)
        assert st.moves_scored == 2 * n * (n - 1) // 2 and st.steps_accepted == 2
        mv, sc, total = e.trace(0)
        r = rows0.copy()
        for (i, j), s in zip(mv, sc):
            r[i], r[j] = r[j], r[i]
            assert orc.nq_score(r) == int(s)
        assert np.array_equal(r, e.get_chains()[0]) and int(sc[-1]) < s0
        rng = np.random.default_rng(0)
        a = rng.integers(0, n - 1, 16)
        b = a + 1 + rng.integers(0, n, 16) % (n - 1 - a)
        assert np.array_equal(e.eval_moves(a, b, 0, cs.SWAP), orc.nq_eval_moves(r, a, b, orc.SWAP))
    with pytest.raises(cs.CsError) as err:
        cs.NQueensChains(n, 2)  # the big path holds one instance per handle
    assert err.value.status == L.CS_ERR_UNSUPPORTED
    with pytest.raises(cs.CsError):
        cs.NQueensChains(n, 1, neighbourhood=cs.CHANGE)


def test_million_queens_partition_properties():
    """BASELINE config 5 size (n = 10^6): one 1/64 slice of the swap neighbourhood, then
    size-independent checks -- the winning key decodes to a move whose delta the scalar hook
    reproduces, and the delta-tracked score equals a fresh counter rebuild."""
    import torch
    from constraint_solver_b200.dist import device_view

    n = 1_000_000
    with cs.NQueensChains(n, 1, seed=42, trace_capacity=2) as e:
        e.init_random()
        rows0 = e.get_chains()[0]
        assert sorted(rows0[:1000].tolist()) != list(range(1000)) and len(np.unique(rows0)) == n
        s0 = int(e.scores()[0])
        e.set_partition(63, 64)
        e.part_scan()
        torch.cuda.synchronize()  # the handle runs on its own stream
        key = int(device_view(e.part_key_device_ptr(), (1,), "<i8", torch.device("cuda", 0)).item())
        v, i, j = (key >> 40) - (1 << 22), (key >> 20) & 0xFFFFF, key & 0xFFFFF
        assert 0 <= i < j < n
        assert int(e.eval_moves([i], [j], 0, cs.SWAP)[0]) == 2 * v
        st = e.part_apply()
        assert st.steps_accepted == 1 and st.best_score == s0 + 2 * v
        assert abs(st.moves_scored - n * (n - 1) // 2 / 64) < n  # triangular-balanced slice
        rows1 = e.get_chains()[0]
        exp = rows0.copy()
        exp[i], exp[j] = exp[j], exp[i]
        assert np.array_equal(rows1, exp)
        with cs.NQueensChains(n, 1) as f:  # fresh rebuild of the counters from the new rows
            f.set_chains(rows1)
            assert int(f.scores()[0]) == s0 + 2 * v


def test_packed_global_scan_equals_oracle_and_scalar_scan():
    """nqb_scan_packed_kernel (byte counters in global memory, 16x2 SIMD, hi/lo diagonal ids) runs
    for permutation boards with n >= 256: every candidate delta == the oracle's full re-score
    difference, and the trajectory equals the scalar global scan (CS_NQ_FLAG_SCALAR) and the
    shared-memory path."""
    rng = np.random.default_rng(11)
    for n in (256, 257, 300, 383, 384, 500, 641):
        rows = np.ascontiguousarray(rng.permutation(n), dtype=np.int64)
        with cs.NQueensChains(n, 1, force_global=True) as e:
            e.set_chains(rows)
            dev = e.neighbourhood_deltas(0)
        ref = orc.nq_neighbourhood_deltas(rows, orc.SWAP)
        bad = np.nonzero(dev != ref)[0]
        assert bad.size == 0, (n, bad[:5], dev[bad[:5]], ref[bad[:5]])
    # many attacking pairs: the identity permutation puts n queens on one diagonal -> a line
    # longer than 62 -> the packed scan must decline and the scalar scan must answer
    for rows in (np.arange(300, dtype=np.int64), np.arange(300, dtype=np.int64)[::-1].copy()):
        with cs.NQueensChains(300, 1, force_global=True) as e:
            e.set_chains(rows)
            assert np.array_equal(e.neighbourhood_deltas(0), orc.nq_neighbourhood_deltas(rows, orc.SWAP))
    # a board with long-but-legal lines (blocks of 40 on one diagonal) stays on the packed scan
    n = 400
    rows = np.arange(n, dtype=np.int64)
    for s in range(0, n, 40):
        rows[s:s + 40] = (rows[s:s + 40] + 7 * (s // 40)) % n
    if len(set(rows.tolist())) == n:
        with cs.NQueensChains(n, 1, force_global=True) as e:
            e.set_chains(rows)
            assert np.array_equal(e.neighbourhood_deltas(0), orc.nq_neighbourhood_deltas(rows, orc.SWAP))
    for n, seed in [(300, 1), (1000, 2), (4000, 3)]:
        start = orc.nq_init_perm(seed, 0, n)
        with cs.NQueensChains(n, 1, trace_capacity=16, force_global=True) as a, \
                cs.NQueensChains(n, 1, trace_capacity=16, force_global=True, force_scalar=True) as b, \
                cs.NQueensChains(n, 1, trace_capacity=16) as c:
            for e in (a, b, c):
                e.set_chains(start)
            sa, sb, sc = a.step(6), b.step(6), c.step(6)
            ma, ca, ta = a.trace(0)
            for e, st in ((b, sb), (c, sc)):
                m2, c2, t2 = e.trace(0)
                assert ta == t2 and np.array_equal(ma, m2) and np.array_equal(ca, c2)
                assert st.moves_scored == sa.moves_scored and st.best_score == sa.best_score
            r = start.copy()
            for (i, j), s in zip(ma, ca):  # replay through the oracle scorer
                r[i], r[j] = r[j], r[i]
                assert orc.nq_score(r) == int(s)


def test_packed_global_scan_short_work_units(monkeypatch):
    """CS_NQB_SEG=3: every column group's j sweep is cut into 3-chunk work units (the multi-GPU
    load-balancing path); deltas and trajectories must not change."""
    monkeypatch.setenv("CS_NQB_SEG", "3")
    rng = np.random.default_rng(12)
    for n in (257, 400, 641):
        rows = np.ascontiguousarray(rng.permutation(n), dtype=np.int64)
        with cs.NQueensChains(n, 1, force_global=True) as e:
            e.set_chains(rows)
            assert np.array_equal(e.neighbourhood_deltas(0), orc.nq_neighbourhood_deltas(rows, orc.SWAP))
    start = orc.nq_init_perm(5, 0, 2000)
    with cs.NQueensChains(2000, 1, trace_capacity=8, force_global=True) as a, \
            cs.NQueensChains(2000, 1, trace_capacity=8) as c:
        a.set_chains(start)
        c.set_chains(start)
        sa, sc = a.step(5), c.step(5)
        assert sa.moves_scored == sc.moves_scored == 5 * 2000 * 1999 // 2
        ma, ca, ta = a.trace(0)
        mc, cc, tc = c.trace(0)
        assert ta == tc and np.array_equal(ma, mc) and np.array_equal(ca, cc)
    monkeypatch.delenv("CS_NQB_SEG")
    _partitions_impl(seg="2")


def test_packed_global_scan_partitions():
    _partitions_impl(seg=None)


def _partitions_impl(seg=None):
    if seg is not None:
        import os
        os.environ["CS_NQB_SEG"] = seg
    try:
        _partitions_body()
    finally:
        if seg is not None:
            os.environ.pop("CS_NQB_SEG", None)


def _partitions_body():
    """5 partitions of a permutation board (packed scan per slice): min of the keys applied on
    every replica == the unpartitioned step; the slices cover the neighbourhood exactly once."""
    import torch
    from constraint_solver_b200.dist import device_view

    n, parts = 3000, 5
    start = orc.nq_init_perm(9, 0, n)
    with cs.NQueensChains(n, 1, trace_capacity=4, force_global=True) as whole:
        whole.set_chains(start)
        whole.step(2)
        wmv, wsc, _ = whole.trace(0)
        want = whole.get_chains()[0]
    engines = [cs.NQueensChains(n, 1, trace_capacity=4, force_global=True) for _ in range(parts)]
    try:
        views = []
        for k, e in enumerate(engines):
            e.set_chains(start)
            e.set_partition(k, parts)
            views.append(device_view(e.part_key_device_ptr(), (1,), "<i8", torch.device("cuda", 0)))
        for step in range(2):
            for e in engines:
                e.part_scan()
            torch.cuda.synchronize()
            best = min(int(v.item()) for v in views)
            for v in views:
                v.fill_(best)
            torch.cuda.synchronize()
            moves = sum(e.part_apply().moves_scored for e in engines)
            assert moves == n * (n - 1) // 2
        for e in engines:
            mv, sc, _ = e.trace(0)
            assert np.array_equal(mv, wmv) and np.array_equal(sc, wsc)
            assert np.array_equal(e.get_chains()[0], want)
    finally:
        for e in engines:
            e.close()
