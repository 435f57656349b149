"""GPU parity at the BENCHMARKED sizes ("what's weak" 1-3): every candidate the
production scans produce is checked, not a sample and not the scalar hook kernel.

The checker for sizes where clone + full re-score of every candidate is out of reach
(5e7 x 5e7 pair tests at n = 10 000) is the oracle's O(1) counter/delta scorer
(oracle/cs_oracle.c: orc_nq_fast_band_deltas), which tests/test_oracle_cpu.py first proves equal to
the literal clone + full re-score on every candidate of small boards, permutations or not.

  * packed shared-memory scan (nq_step_kernel_v2): EVERY entry at n = 10 000 and n = 12 096;
  * packed global scan (nqb_scan_packed_kernel): dumped column bands at n = 20 000, 40 000,
    200 000 and 10^6 (first / middle, tile-unaligned / last columns);
  * the alias-repair pass of the global packed scan on a board built to contain low-15-bit
    diagonal-id aliases inside one tile, plus the chosen move;
  * packed vs CS_NQ_FLAG_SCALAR trajectories at n = 40 000 and 200 000;
  * scheduling 56 x 2000 (configs[3]): every one of the 113 540 candidates vs clone + re-score.
"""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _assert_equal(dev, ref, what):
    if not np.array_equal(dev, ref):
        bad = np.nonzero(dev != ref)[0]
        raise AssertionError((what, int(bad.size), bad[:5].tolist(), dev[bad[:5]].tolist(), ref[bad[:5]].tolist()))


def _long_diagonal_board(n, k, rng):
    rows = np.arange(n)
    tail = rows[k:].copy()
    while True:
        rng.shuffle(tail)
        if not np.any(tail == np.arange(k, n)):
            break
    rows[k:] = tail
    return rows.astype(np.int64)


@pytest.mark.parametrize("n", [10_000, 12_096])
def test_every_candidate_of_the_packed_smem_scan_at_benchmark_size(n):
    rng = np.random.default_rng(n)
    bench_chain = orc.nq_init_perm(42, 0, n)          # the bench's own chain 0 (seed 42)
    boards = [bench_chain, _long_diagonal_board(n, 62, rng)]   # 62 on one line: still the packed path
    if n == 10_000:
        nonperm = rng.permutation(n).astype(np.int64)
        nonperm[17] = nonperm[9000]                    # scalar layout inside the same kernel
        boards += [nonperm, _long_diagonal_board(n, 63, rng)]  # 63: leaves the packed path
    with cs.NQueensChains(n, len(boards), trace_capacity=4) as e:
        e.set_chains(np.stack(boards))
        for k, rows in enumerate(boards):
            dev = e.neighbourhood_deltas(k)
            ref = orc.nq_fast_band_deltas(rows)
            assert dev.size == n * (n - 1) // 2
            _assert_equal(dev, ref, (n, k))
            del dev, ref
        # and after the chain has moved (counters patched in place, not rebuilt): chain 0, 3 steps in
        e.step(3)
        rows3 = e.get_chains()[0]
        _assert_equal(e.neighbourhood_deltas(0), orc.nq_fast_band_deltas(rows3), (n, "after 3 steps"))
        # band == slice of the full dump
        band = e.band_deltas(4999, 5003, 0)
        _assert_equal(band, orc.nq_fast_band_deltas(rows3, 4999, 5003), (n, "band"))


def _bands(n):
    mid = n // 2 - 5  # not a multiple of 16: the band starts and ends inside a tile
    return [(0, 16), (mid, mid + 21), (max(n - 300, 0), n - 1)]


@pytest.mark.parametrize("n", [20_000, 40_000, 200_000, 1_000_000])
def test_column_bands_of_the_packed_global_scan(n):
    with cs.NQueensChains(n, 1, seed=42, trace_capacity=4) as e:
        e.init_random()
        rows = e.get_chains()[0]
        assert np.array_equal(rows, orc.nq_init_perm(42, 0, n))
        for b0, b1 in _bands(n):
            dev = e.band_deltas(b0, b1)
            ref = orc.nq_fast_band_deltas(rows, b0, b1)
            assert dev.size == orc.nq_band_size(n, b0, b1)
            _assert_equal(dev, ref, (n, b0, b1))
        # the selected move is the true (delta, i, j) minimum of the WHOLE neighbourhood
        if n <= 200_000:
            d, a, b, scored = orc.nq_fast_argmin(rows)
            s0 = int(e.scores()[0])
            st = e.step(1)
            mv, sc, _ = e.trace(0)
            assert (int(mv[0][0]), int(mv[0][1])) == (a, b) and int(sc[0]) == s0 + d
            assert st.moves_scored == scored
    # the scalar global scan (CS_NQ_FLAG_SCALAR) on the same bands
    if n <= 40_000:
        with cs.NQueensChains(n, 1, seed=42, force_scalar=True) as e:
            e.init_random()
            rows = e.get_chains()[0]
            for b0, b1 in _bands(n)[1:]:
                _assert_equal(e.band_deltas(b0, b1), orc.nq_fast_band_deltas(rows, b0, b1), (n, "scalar", b0))


def _alias_board(n, i0, rng):
    """Permutation with forced low-15-bit diagonal-id aliases inside the tile [i0, i0+16): for
    tile column i and a partner j, D1 ids r-c+n differ by exactly 2^15 (and D2 ids r+c likewise)
    so the packed scan's 15-bit attack test matches although the queens do not attack; the
    exact repair pass must take the 2 back.  Few partners per line, so no line exceeds 62."""
    rows = rng.permutation(n).astype(np.int64)
    pos = np.empty(n, dtype=np.int64)
    pos[rows] = np.arange(n)
    protected = set(range(i0, i0 + 16))
    forced = []

    def place(j, target):  # rows[j] := target by a swap that keeps the permutation
        if not (0 <= target < n) or j in protected:
            return False
        p = int(pos[target])
        if p in protected or p == j:
            return False
        rj = int(rows[j])
        rows[j], rows[p] = target, rj
        pos[target], pos[rj] = j, p
        protected.add(j)
        protected.add(p)
        return True

    for t in range(16):
        i = i0 + t
        ri = int(rows[i])
        for s, k in ((+1, 1), (-1, 1)):
            for rep in range(3):
                j = int(rng.integers(i0 + 16, n))
                # D1 alias: (rj - j) = (ri - i) + s * 32768
                if place(j, j + (ri - i) + s * 32768):
                    forced.append((i, j, "d1"))
                j = int(rng.integers(i0 + 16, n))
                # D2 alias: (rj + j) = (ri + i) + s * 32768
                if place(j, (ri + i) + s * 32768 - j):
                    forced.append((i, j, "d2"))
    return rows, forced


def test_alias_repair_pass_of_the_packed_global_scan():
    n, i0 = 70_000, 33_008  # i0 multiple of 16 = one warp tile of the packed global scan
    rng = np.random.default_rng(2025)
    rows, forced = _alias_board(n, i0, rng)
    assert sorted(rows.tolist()) == list(range(n))
    real = [(i, j, w) for (i, j, w) in forced
            if (w == "d1" and abs((rows[j] - j) - (rows[i] - i)) == 32768)
            or (w == "d2" and abs((rows[j] + j) - (rows[i] + i)) == 32768)]
    assert len(real) >= 40, len(real)   # the tile really holds aliased (i, j) pairs
    with cs.NQueensChains(n, 1, trace_capacity=4) as e:
        e.set_chains(rows)
        dev = e.band_deltas(i0, i0 + 16)
        ref = orc.nq_fast_band_deltas(rows, i0, i0 + 16)
        _assert_equal(dev, ref, "alias band")
        # the aliased candidates themselves
        tri = lambda x: x * n - x * (x + 1) // 2
        for i, j, _ in real:
            k = tri(i) - tri(i0) + (j - i - 1)
            assert dev[k] == ref[k]
        d, a, b, scored = orc.nq_fast_argmin(rows)
        s0 = int(e.scores()[0])
        st = e.step(1)
        mv, sc, _ = e.trace(0)
        assert (int(mv[0][0]), int(mv[0][1])) == (a, b) and int(sc[0]) == s0 + d and st.moves_scored == scored


@pytest.mark.parametrize("n,steps", [(40_000, 3), (200_000, 2)])
def test_packed_and_scalar_global_scans_walk_the_same_trajectory(n, steps):
    with cs.NQueensChains(n, 1, seed=9, trace_capacity=8) as a, \
            cs.NQueensChains(n, 1, seed=9, trace_capacity=8, force_scalar=True) as b:
        a.init_random()
        b.init_random()
        sa, sb = a.step(steps), b.step(steps)
        ma, ca, ta = a.trace(0)
        mb, cb, tb = b.trace(0)
        assert ta == tb == steps and np.array_equal(ma, mb) and np.array_equal(ca, cb)
        assert sa.moves_scored == sb.moves_scored == steps * n * (n - 1) // 2
        assert np.array_equal(a.get_chains(), b.get_chains())


def test_out_of_range_rows_are_rejected_and_never_indexed():
    """-1 / 65535 used to reach the counter build (shared-memory index far out of
    bounds).  Now: flagged, stored as 0, the handle stays usable."""
    n = 300
    good = orc.nq_init_perm(3, 0, n)
    with cs.NQueensChains(n, 2, trace_capacity=4) as e:
        e.set_chains(np.stack([good, good]))
        for badval in (-1, 65535, n, 2**40):
            bad = good.copy()
            bad[7] = badval
            with pytest.raises(cs.CsError) as err:
                e.set_chains(bad, first_chain=1)
            assert err.value.status == L.CS_ERR_INVALID_ARG
            stored = e.get_chains()[1]
            expect = good.copy()
            expect[7] = 0
            assert np.array_equal(stored, expect)
            assert int(e.scores()[1]) == orc.nq_score(expect) == e.score_full(1)
            e.step(1)  # still healthy
            e.set_chains(good, first_chain=1)
        # the u16 device setter validates too
        import torch
        dev_rows = torch.from_numpy(good.astype(np.int16)).cuda()
        dev_rows[11] = -1  # 65535 as u16
        with pytest.raises(cs.CsError):
            e.set_chain_from_device(0, dev_rows.data_ptr())
        expect = good.copy()
        expect[11] = 0
        assert np.array_equal(e.get_chains()[0], expect) and int(e.scores()[0]) == orc.nq_score(expect)
    with cs.NQueensChains(20_000, 1) as e:  # big-board packer
        bad = orc.nq_init_perm(3, 0, 20_000)
        bad[5] = -1
        with pytest.raises(cs.CsError):
            e.set_chains(bad)
        bad[5] = 0
        assert np.array_equal(e.get_chains()[0], bad) and int(e.scores()[0]) == e.score_full(0)


def test_handles_of_different_sizes_coexist():
    """the dynamic shared-memory opt-in is per function, not per handle."""
    big = cs.NQueensChains(10_000, 2, trace_capacity=2)
    try:
        big.init_random()
        with cs.NQueensChains(300, 2) as small:   # used to lower the limit of the live big handle
            small.init_random()
            small.step(1)
        with cs.ScheduleChains(56, np.arange(2000), n_chains=4) as es_big:
            es_big.init_random()
            with cs.ScheduleChains(7, np.arange(3), n_chains=4) as es_small:
                es_small.init_random()
                es_small.step(1)
            es_big.step(1)
        st = big.step(1)
        assert st.moves_scored == 2 * 10_000 * 9_999 // 2
    finally:
        big.close()


def _es2000():
    rng = np.random.default_rng(42)
    D, E = 56, 2000
    ids = np.arange(E)
    hol = [(int(e), int(d)) for e in range(E) for d in rng.choice(D, size=4, replace=False)]
    return D, E, ids, hol


def test_every_candidate_of_scheduling_56x2000():
    D, E, ids, hol = _es2000()
    rng = np.random.default_rng(1)
    starts = [orc.es_init(42, 0, D + 1, ids),                        # the bench's own chain 0
              ids[rng.integers(0, 12, size=D + 1)],                  # few employees: long masks, windows over cap
              ids[(np.arange(D + 1) * 37) % E]]                      # everyone at most once
    with cs.ScheduleChains(D, ids, holidays=hol, n_chains=len(starts), trace_capacity=8) as e:
        e.set_chains(np.stack(starts))
        for k, a in enumerate(starts):
            dev_h, dev_s = e.neighbourhood_deltas(k)
            ref_h, ref_s = orc.es_neighbourhood_deltas(a[:D], ids, 0, hol)
            assert dev_h.size == D * E + D * (D - 1) // 2 == 113_540
            _assert_equal(dev_h, ref_h, ("hard", k))
            _assert_equal(dev_s, ref_s, ("soft", k))
        e.step(4)  # and from a state the device itself produced
        a = e.get_chains()[0]
        dev_h, dev_s = e.neighbourhood_deltas(0)
        ref_h, ref_s = orc.es_neighbourhood_deltas(a[:D], ids, 0, hol)
        _assert_equal(dev_h, ref_h, "hard after 4 steps")
        _assert_equal(dev_s, ref_s, "soft after 4 steps")
