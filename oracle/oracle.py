"""ctypes loader + numpy wrappers for oracle/libcs_oracle.so.

TEST INFRASTRUCTURE ONLY: the CPU restatement of the reference's algorithm
(oracle/cs_oracle.c cites the reference file:line for each function).  The product
package (constraint_solver_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcs_oracle.so")

I64 = C.c_int64
U64 = C.c_uint64
P64 = C.POINTER(C.c_int64)
INT64_MAX = np.iinfo(np.int64).max


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
        os.path.join(_HERE, "cs_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_nq_score.restype = I64
        _lib.orc_nq_neighbourhood_deltas.restype = I64
        _lib.orc_nq_local_search.restype = I64
        _lib.orc_nq_baseline_sample.restype = I64
        _lib.orc_es_local_search.restype = I64
        _lib.orc_es_baseline_sample.restype = I64
        _lib.orc_days_from_civil.restype = I64
        _lib.orc_philox_draw.restype = C.c_uint32
        _lib.orc_nq_ils.restype = I64
        _lib.orc_nq_local_search_ref.restype = I64
        _lib.orc_es_ils.restype = I64
        _lib.orc_es_local_search_ref.restype = I64
        _lib.orc_es_ils_ref.restype = I64
        _lib.orc_nq_fast_band_deltas.restype = I64
        _lib.orc_nq_fast_argmin.restype = I64
        _lib.orc_nq_delta_baseline.restype = I64
        _lib.orc_nq_neighbourhood_deltas_mt.restype = I64
        _lib.orc_es_neighbourhood_deltas_mt.restype = I64
        _lib.orc_esx_neighbourhood_deltas_mt.restype = I64
        _lib.orc_esx_local_search.restype = I64
        _lib.orc_esx_baseline_sample.restype = I64
    return _lib


def _p(a):
    return a.ctypes.data_as(P64) if a is not None else None


def _i64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64))


# ---------------------------------------------------------------- philox
def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def philox_stream(seed, chain, purpose, counter):
    o = (C.c_uint32 * 4)()
    lib().orc_philox_stream(U64(seed), C.c_uint32(chain), C.c_uint32(purpose), U64(counter), o)
    return [int(x) for x in o]


# ---------------------------------------------------------------- n-queens
def nq_col_scores(rows):
    r = _i64(rows)
    out = np.zeros(len(r), dtype=np.int64)
    lib().orc_nq_col_scores(_p(r), I64(len(r)), _p(out))
    return out


def nq_score(rows) -> int:
    r = _i64(rows)
    return int(lib().orc_nq_score(_p(r), I64(len(r))))


def nq_init_perm(seed, chain, n):
    out = np.zeros(n, dtype=np.int64)
    lib().orc_nq_init_perm(U64(seed), C.c_uint32(chain), I64(n), _p(out))
    return out


SWAP, CHANGE = 0, 1
TIE_MOVE_ORDER, TIE_REFERENCE = 0, 1


def nq_neighbourhood_deltas(rows, kind=SWAP):
    r = _i64(rows)
    n = len(r)
    cnt = n * (n - 1) // 2 if kind == SWAP else n * n
    out = np.zeros(max(cnt, 1), dtype=np.int64)
    k = lib().orc_nq_neighbourhood_deltas(_p(r), I64(n), C.c_int(kind), _p(out))
    assert k == cnt
    return out[:cnt]


def nq_eval_moves(rows, a, b, kind=SWAP):
    r, a, b = _i64(rows), _i64(a), _i64(b)
    out = np.zeros(max(len(a), 1), dtype=np.int64)
    lib().orc_nq_eval_moves(_p(r), I64(len(r)), C.c_int(kind), _p(a), _p(b), I64(len(a)), _p(out))
    return out[: len(a)]


def nq_local_search(rows, kind=SWAP, tie=TIE_MOVE_ORDER, allow_no_improvement_for=5,
                    max_iterations=10_000, window_size=0, trace_cap=0):
    """LocalSearch::execute restated (local_search.rs:301-342). Returns a dict."""
    r = _i64(rows).copy()
    n = len(r)
    cur = np.zeros(max(n, 1), dtype=np.int64)
    best_score, cur_score = I64(0), I64(0)
    ta = np.zeros(max(trace_cap, 1), dtype=np.int64)
    tb = np.zeros(max(trace_cap, 1), dtype=np.int64)
    ts = np.zeros(max(trace_cap, 1), dtype=np.int64)
    steps = lib().orc_nq_local_search(
        _p(r), I64(n), C.c_int(kind), C.c_int(tie), U64(allow_no_improvement_for),
        U64(max_iterations), U64(window_size), C.byref(best_score), _p(cur),
        C.byref(cur_score), _p(ta), _p(tb), _p(ts), I64(trace_cap))
    k = min(int(steps), trace_cap)
    return dict(best=r, best_score=int(best_score.value), current=cur[:n],
                current_score=int(cur_score.value), steps=int(steps),
                trace_a=ta[:k], trace_b=tb[:k], trace_score=ts[:k])


def nq_baseline_sample(rows, a, b, threads):
    r, a, b = _i64(rows), _i64(a), _i64(b)
    chk = I64(0)
    k = lib().orc_nq_baseline_sample(_p(r), I64(len(r)), _p(a), _p(b), I64(len(a)),
                                     C.c_int(threads), C.byref(chk))
    return int(k), int(chk.value)


def _threads(threads):
    return int(threads) if threads else (os.cpu_count() or 1)


def nq_band_size(n, x_begin, x_end, kind=SWAP):
    if kind == SWAP:
        tri = lambda x: x * n - x * (x + 1) // 2
        return tri(x_end) - tri(x_begin)
    return (x_end - x_begin) * n


def nq_fast_band_deltas(rows, x_begin=0, x_end=None, kind=SWAP, threads=None):
    """Checker-side O(1) delta scorer (counters + SURVEY 8(a2) formulae), band-relative order."""
    r = _i64(rows)
    n = len(r)
    x_end = (n - 1 if kind == SWAP else n) if x_end is None else x_end
    cnt = nq_band_size(n, x_begin, x_end, kind)
    out = np.empty(max(cnt, 1), dtype=np.int64)
    k = lib().orc_nq_fast_band_deltas(_p(r), I64(n), C.c_int(kind), I64(x_begin), I64(x_end),
                                      C.c_int(_threads(threads)), _p(out))
    assert k == cnt, (k, cnt)
    return out[:cnt]


def nq_fast_argmin(rows, kind=SWAP, threads=None):
    """(delta, a, b, candidates scored) of the whole neighbourhood by the (delta, a, b) rule."""
    r = _i64(rows)
    d, a, b = I64(0), I64(0), I64(0)
    k = lib().orc_nq_fast_argmin(_p(r), I64(len(r)), C.c_int(kind), C.c_int(_threads(threads)),
                                 C.byref(d), C.byref(a), C.byref(b))
    return int(d.value), int(a.value), int(b.value), int(k)


def nq_delta_baseline(seed, n, chains, steps, threads=None):
    chk = I64(0)
    k = lib().orc_nq_delta_baseline(U64(seed), I64(n), C.c_int(chains), C.c_int(steps),
                                    C.c_int(_threads(threads)), C.byref(chk))
    return int(k), int(chk.value)


def nq_neighbourhood_deltas_mt(rows, kind=SWAP, threads=None):
    r = _i64(rows)
    n = len(r)
    cnt = n * (n - 1) // 2 if kind == SWAP else n * n
    out = np.zeros(max(cnt, 1), dtype=np.int64)
    k = lib().orc_nq_neighbourhood_deltas_mt(_p(r), I64(n), C.c_int(kind), C.c_int(_threads(threads)), _p(out))
    assert k == cnt
    return out[:cnt]


# ---------------------------------------------------------------- employee scheduling
def days_from_civil(y, m, d) -> int:
    return int(lib().orc_days_from_civil(I64(y), C.c_int(m), C.c_int(d)))


def weekday_from_days(z) -> int:
    return int(lib().orc_weekday_from_days(I64(z)))


def weekday(y, m, d) -> int:
    return weekday_from_days(days_from_civil(y, m, d))


def _hol(holidays):
    if holidays is None or len(holidays) == 0:
        return _i64([]), _i64([])
    h = _i64(holidays).reshape(-1, 2)
    return np.ascontiguousarray(h[:, 0]), np.ascontiguousarray(h[:, 1])


def es_score_terms(a, start_weekday=0, holidays=None):
    """holidays: iterable of (employee_id, day_index). Returns int64[8] = H1..H4,S1..S4."""
    a = _i64(a)
    he, hd = _hol(holidays)
    out = np.zeros(8, dtype=np.int64)
    rc = lib().orc_es_score_terms(_p(a), I64(len(a)), C.c_int(start_weekday), _p(he), _p(hd),
                                  I64(len(he)), _p(out))
    if rc:
        raise ValueError("holiday outside the scored range (the reference panics: lib.rs:275)")
    return out


def es_score(a, start_weekday=0, holidays=None):
    t = es_score_terms(a, start_weekday, holidays)
    return int(t[:4].sum()), int(t[4:].sum())


ES_CHANGE, ES_SWAP = 0, 1


def es_init(seed, chain, n_slots, employees):
    employees = _i64(employees)
    out = np.zeros(n_slots, dtype=np.int64)
    lib().orc_es_init(U64(seed), C.c_uint32(chain), I64(n_slots), _p(employees),
                      I64(len(employees)), _p(out))
    return out


def es_eval_moves(a, employees, x, y, kind, start_weekday=0, holidays=None):
    a, employees, x, y = _i64(a), _i64(employees), _i64(x), _i64(y)
    he, hd = _hol(holidays)
    dh = np.zeros(max(len(x), 1), dtype=np.int64)
    ds = np.zeros(max(len(x), 1), dtype=np.int64)
    rc = lib().orc_es_eval_moves(_p(a), I64(len(a)), C.c_int(start_weekday), _p(he), _p(hd),
                                 I64(len(he)), _p(employees), I64(len(employees)),
                                 C.c_int(kind), _p(x), _p(y), I64(len(x)), _p(dh), _p(ds))
    if rc:
        raise ValueError("holiday outside the scored range")
    return dh[: len(x)], ds[: len(x)]


def es_neighbourhood_deltas(a, employees, start_weekday=0, holidays=None, threads=None):
    """(dhard, dsoft) of every candidate, device enumeration order, identities = INT64_MAX."""
    a, employees = _i64(a), _i64(employees)
    he, hd = _hol(holidays)
    D, E = len(a), len(employees)
    cnt = D * E + D * (D - 1) // 2
    dh = np.zeros(max(cnt, 1), dtype=np.int64)
    ds = np.zeros(max(cnt, 1), dtype=np.int64)
    k = lib().orc_es_neighbourhood_deltas_mt(_p(a), I64(D), C.c_int(start_weekday), _p(he), _p(hd), I64(len(he)),
                                             _p(employees), I64(E), C.c_int(_threads(threads)), _p(dh), _p(ds))
    if k < 0:
        raise ValueError("holiday outside the scored range")
    assert k == cnt
    return dh[:cnt], ds[:cnt]


def es_local_search(a, employees, start_weekday=0, holidays=None,
                    allow_no_improvement_for=20, max_iterations=1000, trace_cap=0):
    a = _i64(a).copy()
    employees = _i64(employees)
    he, hd = _hol(holidays)
    D = len(a)
    cur = np.zeros(max(D, 1), dtype=np.int64)
    bh, bs = I64(0), I64(0)
    tr = [np.zeros(max(trace_cap, 1), dtype=np.int64) for _ in range(5)]
    steps = lib().orc_es_local_search(
        _p(a), I64(D), C.c_int(start_weekday), _p(he), _p(hd), I64(len(he)), _p(employees),
        I64(len(employees)), U64(allow_no_improvement_for), U64(max_iterations),
        C.byref(bh), C.byref(bs), _p(cur), _p(tr[0]), _p(tr[1]), _p(tr[2]), _p(tr[3]),
        _p(tr[4]), I64(trace_cap))
    k = min(int(steps), trace_cap)
    return dict(best=a, best_hard=int(bh.value), best_soft=int(bs.value), current=cur[:D],
                steps=int(steps), trace_kind=tr[0][:k], trace_x=tr[1][:k], trace_y=tr[2][:k],
                trace_hard=tr[3][:k], trace_soft=tr[4][:k])


def es_baseline_sample(a, employees, x, y, kind, threads, start_weekday=0, holidays=None):
    a, employees, x, y = _i64(a), _i64(employees), _i64(x), _i64(y)
    he, hd = _hol(holidays)
    chk = I64(0)
    k = lib().orc_es_baseline_sample(_p(a), I64(len(a)), C.c_int(start_weekday), _p(he),
                                     _p(hd), I64(len(he)), _p(employees), I64(len(employees)),
                                     C.c_int(kind), _p(x), _p(y), I64(len(x)),
                                     C.c_int(threads), C.byref(chk))
    return int(k), int(chk.value)


# ---------------------------------------------------------------- iterated local search
def nq_local_search_ref(rows, seed, chain, rng_t=0, allow_no_improvement_for=5,
                        max_iterations=10_000, window_size=None, trace_cap=0):
    """LocalSearch::execute with the reference's own proposer / window / tie-break."""
    r = _i64(rows).copy()
    n = len(r)
    window_size = 5 * n if window_size is None else window_size
    cur = np.zeros(max(n, 1), dtype=np.int64)
    best_score = I64(0)
    t = U64(rng_t)
    ta = np.zeros(max(trace_cap, 1), dtype=np.int64)
    tb = np.zeros(max(trace_cap, 1), dtype=np.int64)
    ts = np.zeros(max(trace_cap, 1), dtype=np.int64)
    steps = lib().orc_nq_local_search_ref(_p(r), I64(n), U64(seed), C.c_uint32(chain), C.byref(t),
                                          U64(allow_no_improvement_for), U64(max_iterations),
                                          U64(window_size), C.byref(best_score), _p(cur), _p(ta),
                                          _p(tb), _p(ts), I64(trace_cap))
    k = min(int(steps), trace_cap)
    return dict(best=r, best_score=int(best_score.value), current=cur[:n], steps=int(steps),
                rng_t=int(t.value), trace_a=ta[:k], trace_b=tb[:k], trace_score=ts[:k])


def nq_ils(seed, chain, n, kind=SWAP, ls_max_iterations=10_000, allow_no_improvement_for=5,
           rounds=100, best_cap=32, ref_window=0):
    """IteratedLocalSearch (iterated_local_search.rs:173-202) restated; see cs_oracle.c."""
    best = np.zeros(max(n, 1), dtype=np.int64)
    cur = np.zeros(max(n, 1), dtype=np.int64)
    rn = np.zeros(max(rounds, 1), dtype=np.int64)
    rc = np.zeros(max(rounds, 1), dtype=np.int64)
    bs = I64(0)
    r = lib().orc_nq_ils(U64(seed), C.c_uint32(chain), I64(n), C.c_int(kind), U64(ls_max_iterations),
                         U64(allow_no_improvement_for), U64(rounds), C.c_int(best_cap), _p(best),
                         C.byref(bs), _p(cur), _p(rn), _p(rc), U64(ref_window))
    r = int(r)
    return dict(rounds=r, best=best[:n], best_score=int(bs.value), current=cur[:n],
                round_new_score=rn[:r], round_choice=rc[:r])


def es_ils(seed, chain, D, employees, start_weekday=0, holidays=None, ls_max_iterations=1000,
           allow_no_improvement_for=20, rounds=50, best_cap=64):
    employees = _i64(employees)
    he, hd = _hol(holidays)
    best = np.zeros(D + 1, dtype=np.int64)
    rn = np.zeros(max(rounds, 1), dtype=np.int64)
    rc = np.zeros(max(rounds, 1), dtype=np.int64)
    bh, bs = I64(0), I64(0)
    r = lib().orc_es_ils(U64(seed), C.c_uint32(chain), I64(D), C.c_int(start_weekday), _p(he), _p(hd),
                         I64(len(he)), _p(employees), I64(len(employees)), U64(ls_max_iterations),
                         U64(allow_no_improvement_for), U64(rounds), C.c_int(best_cap), _p(best),
                         C.byref(bh), C.byref(bs), _p(rn), _p(rc))
    r = int(r)
    return dict(rounds=r, best=employees[best], best_hard=int(bh.value), best_soft=int(bs.value),
                round_new_key=rn[:r], round_choice=rc[:r])


def es_local_search_ref(a, employees, seed, chain, start_weekday=0, holidays=None,
                        allow_no_improvement_for=20, max_iterations=1000, window_size=100,
                        max_draws=1 << 16, trace_cap=0):
    """LocalSearch::execute with the reference's own scheduling proposer (random ChangeDay /
    SwapDays stream from a cloned rng), window and derived-Ord tie-break."""
    a = _i64(a).copy()
    employees = _i64(employees)
    he, hd = _hol(holidays)
    D = len(a)
    cur = np.zeros(max(D, 1), dtype=np.int64)
    bh, bs, sc = I64(0), I64(0), I64(0)
    tr = [np.zeros(max(trace_cap, 1), dtype=np.int64) for _ in range(5)]
    steps = lib().orc_es_local_search_ref(
        _p(a), I64(D), C.c_int(start_weekday), _p(he), _p(hd), I64(len(he)), _p(employees),
        I64(len(employees)), U64(seed), C.c_uint32(chain), U64(allow_no_improvement_for),
        U64(max_iterations), U64(window_size), U64(max_draws), C.byref(bh), C.byref(bs), _p(cur),
        _p(tr[0]), _p(tr[1]), _p(tr[2]), _p(tr[3]), _p(tr[4]), I64(trace_cap), C.byref(sc))
    k = min(int(steps), trace_cap)
    return dict(best=a, best_hard=int(bh.value), best_soft=int(bs.value), current=cur[:D],
                steps=int(steps), scored=int(sc.value), trace_kind=tr[0][:k], trace_x=tr[1][:k],
                trace_y=tr[2][:k], trace_hard=tr[3][:k], trace_soft=tr[4][:k])


def es_ils_ref(seed, chain, D, employees, start_weekday=0, holidays=None, ls_max_iterations=1000,
               allow_no_improvement_for=20, rounds=50, best_cap=64, window_size=100, max_draws=1 << 16):
    employees = _i64(employees)
    he, hd = _hol(holidays)
    best = np.zeros(D + 1, dtype=np.int64)
    rn = np.zeros(max(rounds, 1), dtype=np.int64)
    rc = np.zeros(max(rounds, 1), dtype=np.int64)
    bh, bs = I64(0), I64(0)
    r = lib().orc_es_ils_ref(U64(seed), C.c_uint32(chain), I64(D), C.c_int(start_weekday), _p(he), _p(hd),
                             I64(len(he)), _p(employees), I64(len(employees)), U64(ls_max_iterations),
                             U64(allow_no_improvement_for), U64(rounds), C.c_int(best_cap), _p(best),
                             C.byref(bh), C.byref(bs), _p(rn), _p(rc), U64(window_size), U64(max_draws))
    r = int(r)
    return dict(rounds=r, best=employees[best], best_hard=int(bh.value), best_soft=int(bs.value),
                round_new_key=rn[:r], round_choice=rc[:r])


# ---------------------------------------------------------------- slot-generalised scheduling (extension)
def esx_score_terms(a, employees, D, S, start_weekday=0, holidays=None, skills=None):
    """EXTENSION (not pinned by the reference): D days x S shifts; int64[10] = H1..H4, S1..S4, X1, X2."""
    a, employees = _i64(a), _i64(employees)
    he, hd = _hol(holidays)
    sk = _i64(skills) if skills is not None else None
    out = np.zeros(10, dtype=np.int64)
    rc = lib().orc_esx_score_terms(_p(a), I64(D), I64(S), C.c_int(start_weekday), _p(he), _p(hd), I64(len(he)),
                                   _p(employees), I64(len(employees)), _p(sk), _p(out))
    if rc:
        raise ValueError("holiday outside the scored range")
    return out


def esx_score(a, employees, D, S, start_weekday=0, holidays=None, skills=None):
    t = esx_score_terms(a, employees, D, S, start_weekday, holidays, skills)
    return int(t[:4].sum() + t[8] + t[9]), int(t[4:8].sum())


def esx_neighbourhood_deltas(a, employees, D, S, start_weekday=0, holidays=None, skills=None, threads=None):
    a, employees = _i64(a), _i64(employees)
    he, hd = _hol(holidays)
    sk = _i64(skills) if skills is not None else None
    T, E = D * S, len(employees)
    cnt = T * E + T * (T - 1) // 2
    dh = np.zeros(max(cnt, 1), dtype=np.int64)
    ds = np.zeros(max(cnt, 1), dtype=np.int64)
    k = lib().orc_esx_neighbourhood_deltas_mt(_p(a), I64(D), I64(S), C.c_int(start_weekday), _p(he), _p(hd),
                                              I64(len(he)), _p(employees), I64(E), _p(sk),
                                              C.c_int(_threads(threads)), _p(dh), _p(ds))
    if k < 0:
        raise ValueError("holiday outside the scored range")
    assert k == cnt
    return dh[:cnt], ds[:cnt]


def esx_local_search(a, employees, D, S, start_weekday=0, holidays=None, skills=None,
                     allow_no_improvement_for=20, max_iterations=1000, trace_cap=0, threads=None):
    a = _i64(a).copy()
    employees = _i64(employees)
    he, hd = _hol(holidays)
    sk = _i64(skills) if skills is not None else None
    T = D * S
    cur = np.zeros(max(T, 1), dtype=np.int64)
    bh, bs = I64(0), I64(0)
    tr = [np.zeros(max(trace_cap, 1), dtype=np.int64) for _ in range(5)]
    steps = lib().orc_esx_local_search(
        _p(a), I64(D), I64(S), C.c_int(start_weekday), _p(he), _p(hd), I64(len(he)), _p(employees),
        I64(len(employees)), _p(sk), U64(allow_no_improvement_for), U64(max_iterations),
        C.c_int(_threads(threads)), C.byref(bh), C.byref(bs), _p(cur), _p(tr[0]), _p(tr[1]), _p(tr[2]),
        _p(tr[3]), _p(tr[4]), I64(trace_cap))
    k = min(int(steps), trace_cap)
    return dict(best=a, best_hard=int(bh.value), best_soft=int(bs.value), current=cur[:T], steps=int(steps),
                trace_kind=tr[0][:k], trace_x=tr[1][:k], trace_y=tr[2][:k], trace_hard=tr[3][:k],
                trace_soft=tr[4][:k])


def esx_baseline_sample(a, employees, x, y, kind, threads, D, S, start_weekday=0, holidays=None, skills=None):
    a, employees, x, y = _i64(a), _i64(employees), _i64(x), _i64(y)
    he, hd = _hol(holidays)
    sk = _i64(skills) if skills is not None else None
    chk = I64(0)
    k = lib().orc_esx_baseline_sample(_p(a), I64(D), I64(S), C.c_int(start_weekday), _p(he), _p(hd), I64(len(he)),
                                      _p(employees), I64(len(employees)), _p(sk), C.c_int(kind), _p(x), _p(y),
                                      I64(len(x)), C.c_int(threads), C.byref(chk))
    return int(k), int(chk.value)
