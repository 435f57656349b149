"""GPU parity for slot-generalised scheduling: multi-word slot masks.

  * reference-faithful rotas beyond 64 days (one slot per day, any horizon up to 192; the
    reference takes any date range, lib.rs:181-191): every candidate == the oracle's clone + full
    re-score (orc_es_*, the reference restatement), LocalSearch::execute trajectories, reference
    mode;
  * the EXTENSION -- several shifts per day, same-day overlap and skill terms; NOT pinned by the
    reference, defined by its own CPU full-re-score oracle (oracle/cs_oracle.c: esx_terms), which
    reduces to the reference restatement at one shift per day (tests/test_oracle_cpu.py): every
    candidate, the 10 score terms, trajectories, BASELINE configs[2] / [3] at 84 / 168 slots.
"""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from constraint_solver_b200 import _lib as L
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _states(rng, T, ids):
    E = len(ids)
    out = [ids[rng.integers(0, E, size=T + 1)] for _ in range(2)]
    out.append(np.full(T + 1, ids[0]))                           # one employee holds every slot
    near = np.full(T + 1, ids[0])
    near[T // 2] = ids[E - 1]
    out.append(near)                                             # one slot away from that state
    out.append(ids[np.arange(T + 1) % E])                        # round robin
    out.append(ids[(np.arange(T + 1) // 3) % E])                 # runs of three
    out.append(ids[rng.integers(0, min(E, 3), size=T + 1)])      # crowded: long masks, windows over the caps
    return [np.ascontiguousarray(x, dtype=np.int64) for x in out]


def _check(dev, ref, what):
    for name, d, r in (("hard", dev[0], ref[0]), ("soft", dev[1], ref[1])):
        if not np.array_equal(d, r):
            bad = np.nonzero(d != r)[0]
            raise AssertionError((what, name, int(bad.size), bad[:6].tolist(), d[bad[:6]].tolist(), r[bad[:6]].tolist()))


@pytest.mark.parametrize("D,E,wd", [(65, 5, 0), (90, 11, 3), (128, 4, 6), (129, 30, 2), (168, 7, 5), (192, 3, 1),
                                    (192, 200, 0)])
def test_long_horizons_every_delta_and_trajectory(D, E, wd):
    rng = np.random.default_rng(D * 1000 + E)
    ids = np.sort(rng.choice(np.arange(0, 4 * E + 5), size=E, replace=False)).astype(np.int64)
    hol = sorted({(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(int(rng.integers(0, 3 * E + 1)))})
    states = _states(rng, D, ids)
    with cs.ScheduleChains(D, ids, start_weekday=wd, holidays=hol, n_chains=len(states), trace_capacity=8) as e:
        e.set_chains(np.stack(states))
        hard, soft = e.scores()
        for k, a in enumerate(states):
            assert orc.es_score(a[:D], wd, hol) == (int(hard[k]), int(soft[k])), (D, E, k)
            h, s, terms = e.score_full(k)
            assert (h, s) == (int(hard[k]), int(soft[k])) and terms == orc.es_score_terms(a[:D], wd, hol).tolist()
            _check(e.neighbourhood_deltas(k), orc.es_neighbourhood_deltas(a[:D], ids, wd, hol), (D, E, k))
        e.local_search(3, 5)
        for k in (0, 3, 6):
            ref = orc.es_local_search(states[k][:D], ids, wd, hol, allow_no_improvement_for=3, max_iterations=5,
                                      trace_cap=8)
            mv, th, ts, total = e.trace(k)
            assert total == ref["steps"], (D, E, k)
            assert np.array_equal(mv["kind"], ref["trace_kind"]) and np.array_equal(mv["a"], ref["trace_x"])
            assert np.array_equal(mv["b"], ref["trace_y"])
            assert np.array_equal(th, ref["trace_hard"]) and np.array_equal(ts, ref["trace_soft"])
        best, bh, bs = e.get_best_chains()
        ref = orc.es_local_search(states[0][:D], ids, wd, hol, allow_no_improvement_for=3, max_iterations=5)
        assert (int(bh[0]), int(bs[0])) == (ref["best_hard"], ref["best_soft"]) and np.array_equal(best[0][:D], ref["best"])


def test_long_horizon_reference_mode_and_ils():
    D, E, wd, window = 100, 9, 4, 60
    ids = np.arange(E, dtype=np.int64) * 3 + 1
    hol = [(1, 5), (4, 50), (25, 99)]
    with cs.ScheduleChains(D, ids, start_weekday=wd, holidays=hol, n_chains=3, seed=11, trace_capacity=12,
                           reference_proposer=True) as e:
        e.set_window(window)
        e.init_random()
        start = e.get_chains()
        e.local_search(4, 10)
        for k in range(3):
            ref = orc.es_local_search_ref(start[k][:D], ids, 11, k, wd, hol, 4, 10, window, trace_cap=12)
            mv, th, ts, total = e.trace(k)
            assert total == ref["steps"]
            assert np.array_equal(mv["kind"], ref["trace_kind"]) and np.array_equal(mv["a"], ref["trace_x"])
            assert np.array_equal(mv["b"], ref["trace_y"]) and np.array_equal(th, ref["trace_hard"])
            assert np.array_equal(ts, ref["trace_soft"])
    D, E = 70, 4
    ids = np.arange(E, dtype=np.int64)
    with cs.ScheduleChains(D, ids, holidays=[(0, 3)], n_chains=2, seed=5) as e:
        e.init_random()
        e.ils_init(8, 4)
        e.ils_run(3, 4, 2)
        for k in range(2):
            ref = orc.es_ils(5, k, D, ids, 0, [(0, 3)], ls_max_iterations=4, allow_no_improvement_for=2, rounds=3,
                             best_cap=8)
            key, choice, total = e.ils_log(k)
            assert total == ref["rounds"]
            assert np.array_equal(key, ref["round_new_key"]) and np.array_equal(choice, ref["round_choice"])
            rows, bh, bs = e.ils_best(k)
            assert (bh, bs) == (ref["best_hard"], ref["best_soft"]) and np.array_equal(rows, ref["best"])


SHIFT_CASES = [(1, 2, 2, 0), (2, 3, 3, 5), (7, 3, 4, 0), (10, 2, 5, 3), (21, 3, 6, 6), (14, 3, 40, 2),
               (28, 3, 50, 0), (40, 3, 9, 4), (56, 3, 30, 0), (64, 3, 7, 1), (96, 2, 5, 5), (30, 1, 6, 2)]


@pytest.mark.parametrize("D,S,E,wd", SHIFT_CASES)
def test_shifts_every_delta_terms_and_trajectory(D, S, E, wd):
    rng = np.random.default_rng(D * 100 + S * 10 + E)
    T = D * S
    ids = np.sort(rng.choice(np.arange(0, 4 * E + 5), size=E, replace=False)).astype(np.int64)
    ids_in = ids[rng.permutation(E)]                             # the caller's order is arbitrary
    skills = np.array([int(rng.integers(1, 1 << S)) if rng.random() < 0.6 else (1 << S) - 1 for _ in range(E)])
    if S == 1:
        skills[0] = 0                                            # a skill gap at one shift per day
    hol = [(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(int(rng.integers(0, 2 * E + 1)))]
    states = _states(rng, T, ids)
    sk_sorted = _skills_sorted(ids, ids_in, skills)              # the oracle gets (sorted ids, their skills)
    with cs.ScheduleChains(D, ids_in, start_weekday=wd, holidays=hol, n_chains=len(states), trace_capacity=8,
                           shifts_per_day=S, skills=skills) as e:
        assert e.n_slots == T + 1
        e.set_chains(np.stack(states))
        hard, soft = e.scores()
        for k, a in enumerate(states):
            want = orc.esx_score_terms(a[:T], ids, D, S, wd, hol, sk_sorted)
            assert orc.esx_score(a[:T], ids, D, S, wd, hol, sk_sorted) == (int(hard[k]), int(soft[k])), (k, want)
            h, s, terms = e.score_full_ex(k)
            assert (h, s) == (int(hard[k]), int(soft[k])) and terms == want.tolist(), (k, terms, want.tolist())
            _check(e.neighbourhood_deltas(k), orc.esx_neighbourhood_deltas(a[:T], ids, D, S, wd, hol, sk_sorted),
                   (D, S, E, k))
        e.local_search(3, 4)
        for k in (0, 3, 6):
            ref = orc.esx_local_search(states[k][:T], ids, D, S, wd, hol, sk_sorted, allow_no_improvement_for=3,
                                       max_iterations=4, trace_cap=8)
            mv, th, ts, total = e.trace(k)
            assert total == ref["steps"], (D, S, E, k)
            assert np.array_equal(mv["kind"], ref["trace_kind"]) and np.array_equal(mv["a"], ref["trace_x"])
            assert np.array_equal(mv["b"], ref["trace_y"])
            assert np.array_equal(th, ref["trace_hard"]) and np.array_equal(ts, ref["trace_soft"])


def _skills_sorted(ids_sorted, ids_in, skills):
    pos = {int(v): k for k, v in enumerate(ids_in)}
    return np.array([skills[pos[int(v)]] for v in ids_sorted])


def test_one_shift_per_day_through_the_extension_entry_point_is_the_reference_path():
    """cs_es_create_ex(shifts = 1, everybody qualified) takes the same kernels as cs_es_create: same
    deltas, same trajectory (the extension reduces to the reference's rota)."""
    rng = np.random.default_rng(8)
    D, E = 31, 7
    ids = np.arange(E, dtype=np.int64)
    hol = [(0, 0), (0, 1), (3, 3)]
    start = ids[rng.integers(0, E, size=D + 1)]
    with cs.ScheduleChains(D, ids, holidays=hol, trace_capacity=8) as a, \
            cs.ScheduleChains(D, ids, holidays=hol, trace_capacity=8, shifts_per_day=1, skills=[1] * E) as b:
        a.set_chains(start)
        b.set_chains(start)
        _check(b.neighbourhood_deltas(0), a.neighbourhood_deltas(0), "ex entry point")
        a.local_search(3, 8)
        b.local_search(3, 8)
        ta, tb = a.trace(0), b.trace(0)
        assert all(np.array_equal(x, y) for x, y in zip(ta[:3], tb[:3])) and ta[3] == tb[3]
        assert b.score_full_ex(0)[2][8:] == [0, 0]


@pytest.mark.parametrize("name,D,E,nhol,steps", [("es50x3", 28, 50, 2, 6), ("es2000x3", 56, 2000, 4, 3)])
def test_baseline_configs_at_three_shifts_per_day(name, D, E, nhol, steps):
    """BASELINE configs[2] / [3] as worded ("3 shifts/day": 84 / 168 slots): every candidate of the
    bench's own instance vs the extended oracle, and a replay of device steps."""
    S, T = 3, 3 * D
    rng = np.random.default_rng(42)
    ids = np.arange(E)
    hol = [(int(e), int(d)) for e in range(E) for d in rng.choice(D, size=nhol, replace=False)]
    skills = np.zeros(E, dtype=np.int64)
    for e_ in range(E):
        for s in rng.choice(S, size=2, replace=False):
            skills[e_] |= 1 << int(s)
    with cs.ScheduleChains(D, ids, holidays=hol, n_chains=4, seed=42, trace_capacity=8, shifts_per_day=S,
                           skills=skills) as e:
        e.init_random()
        rows0 = e.get_chains()
        assert np.array_equal(rows0[0], orc.es_init(42, 0, T + 1, ids))
        hard, soft = e.scores()
        assert orc.esx_score(rows0[0][:T], ids, D, S, 0, hol, skills) == (int(hard[0]), int(soft[0]))
        _check(e.neighbourhood_deltas(0), orc.esx_neighbourhood_deltas(rows0[0][:T], ids, D, S, 0, hol, skills), name)
        st = e.step(steps)
        assert st.steps_accepted == 4 * steps
        rows1 = e.get_chains()
        hard, soft = e.scores()
        for k in range(4):
            mv, th, ts, _ = e.trace(k)
            a = rows0[k][:T].copy()
            for m, h, s in zip(mv, th, ts):
                if m["kind"] == 0:
                    a[m["a"]] = ids[m["b"]]
                else:
                    a[m["a"]], a[m["b"]] = a[m["b"]], a[m["a"]]
                assert orc.esx_score(a, ids, D, S, 0, hol, skills) == (int(h), int(s))
            assert np.array_equal(a, rows1[k][:T]) and e.score_full_ex(k)[:2] == (int(hard[k]), int(soft[k]))
        # and from the state the device produced
        _check(e.neighbourhood_deltas(1), orc.esx_neighbourhood_deltas(rows1[1][:T], ids, D, S, 0, hol, skills),
               name + " after steps")


def test_extension_limits_and_errors():
    with pytest.raises(cs.CsError) as err:
        cs.ScheduleChains(10, [0, 1], shifts_per_day=4)
    assert err.value.status == L.CS_ERR_UNSUPPORTED
    with pytest.raises(cs.CsError) as err:
        cs.ScheduleChains(65, [0, 1], shifts_per_day=3)          # 195 slots
    assert err.value.status == L.CS_ERR_UNSUPPORTED
    with pytest.raises(cs.CsError) as err:
        cs.ScheduleChains(193, [0, 1])
    assert err.value.status == L.CS_ERR_UNSUPPORTED
    with pytest.raises(cs.CsError) as err:                       # the reference's proposer belongs to its own rota
        cs.ScheduleChains(10, [0, 1], shifts_per_day=2, reference_proposer=True)
    assert err.value.status == L.CS_ERR_UNSUPPORTED
    with pytest.raises(ValueError):
        cs.ScheduleChains(10, [0, 1], shifts_per_day=2, skills=[1])
    with cs.ScheduleChains(3, [7], shifts_per_day=2) as e:       # one employee: every move is an identity
        e.set_chains(np.full(7, 7))
        st = e.step(2)
        assert st.steps_accepted == 0 and e.status()[0] == L.CHAIN_EMPTY
        assert e.score_full_ex(0)[2] == orc.esx_score_terms([7] * 6, [7], 3, 2).tolist()


def test_random_slot_instances_every_delta_and_trajectory():
    """Randomised sweep over the whole template space (1-3 mask words, with / without the shift and
    skill extension): every candidate and a LocalSearch::execute trajectory against the oracle."""
    rng = np.random.default_rng(20261019)
    for case in range(36):
        S = int(rng.integers(1, 4))
        D = int(rng.integers(1, 192 // S + 1))
        T = D * S
        E = int(rng.integers(1, 40)) if case % 4 else int(rng.integers(1, 5))
        wd = int(rng.integers(0, 7))
        ids = np.sort(rng.choice(np.arange(0, 4 * E + 5), size=E, replace=False)).astype(np.int64)
        skills = None
        if case % 3 != 0:
            skills = np.array([int(rng.integers(0, 1 << S)) if rng.random() < 0.5 else (1 << S) - 1 for _ in range(E)])
        hol = [(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(int(rng.integers(0, 2 * E + 1)))]
        if case % 2:
            p = rng.dirichlet(np.full(E, 0.3))
            start = ids[rng.choice(E, size=T + 1, p=p)]
        else:
            start = ids[rng.integers(0, E, size=T + 1)]
        with cs.ScheduleChains(D, ids, start_weekday=wd, holidays=hol, n_chains=2, trace_capacity=8,
                               shifts_per_day=S, skills=skills) as e:
            e.set_chains(np.stack([start, start]))
            hard, soft = e.scores()
            assert orc.esx_score(start[:T], ids, D, S, wd, hol, skills) == (int(hard[0]), int(soft[0])), (case, D, S, E)
            assert e.score_full_ex(0)[2] == orc.esx_score_terms(start[:T], ids, D, S, wd, hol, skills).tolist()
            _check(e.neighbourhood_deltas(1), orc.esx_neighbourhood_deltas(start[:T], ids, D, S, wd, hol, skills),
                   (case, D, S, E, wd))
            ref = orc.esx_local_search(start[:T], ids, D, S, wd, hol, skills, allow_no_improvement_for=3,
                                       max_iterations=4, trace_cap=8)
            e.local_search(3, 4)
            mv, th, ts, total = e.trace(0)
            assert total == ref["steps"], (case, D, S, E)
            assert np.array_equal(mv["kind"], ref["trace_kind"]) and np.array_equal(mv["a"], ref["trace_x"])
            assert np.array_equal(mv["b"], ref["trace_y"])
            assert np.array_equal(th, ref["trace_hard"]) and np.array_equal(ts, ref["trace_soft"])


def test_packed_day_sets_equal_plain_sets_up_to_the_38_day_limit(monkeypatch):
    """Rotas of <= 38 days keep the 14- and 7-day window sets in one word (step kernel, PK).  Every
    candidate's delta and a 25-step trajectory must equal the plain-set kernel's (the knob) and the
    oracle's, at the limit itself (38 days: 7-day window start 31 = bit 63), one past it (39: not
    packed) and with several shifts per day."""
    rng = np.random.default_rng(38)
    for D, S, E in [(38, 1, 6), (38, 3, 9), (37, 2, 5), (39, 1, 6), (31, 1, 7), (14, 3, 4), (13, 1, 3), (7, 2, 3)]:
        ids = np.arange(E, dtype=np.int64) * 3 + 1
        hol = sorted({(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(E)})
        kw = dict(holidays=hol, n_chains=6, seed=D * 10 + S, trace_capacity=32, start_weekday=int(rng.integers(0, 7)))
        if S > 1:
            kw.update(shifts_per_day=S, skills=[int(rng.integers(1, 1 << S)) for _ in range(E)])
        out = []
        for knob in (None, "1"):
            if knob is None:
                monkeypatch.delenv("CS_ES_NO_PACKED_DAY_SETS", raising=False)
            else:
                monkeypatch.setenv("CS_ES_NO_PACKED_DAY_SETS", knob)
            with cs.ScheduleChains(D, ids, **kw) as e:
                e.init_random()
                start = e.get_chains()
                dh, ds = e.neighbourhood_deltas(0)
                e.step(25)
                out.append((start, dh, ds, e.get_chains(), e.scores(), [e.trace(k) for k in range(6)]))
        a, b = out
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), (D, S)
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4][0], b[4][0]) and np.array_equal(a[4][1], b[4][1])
        for ta, tb in zip(a[5], b[5]):
            assert all(np.array_equal(x, y) for x, y in zip(ta[:3], tb[:3])), (D, S)
        T = D * S
        if S == 1:
            rh, rs = orc.es_neighbourhood_deltas(a[0][0][:T], ids, kw["start_weekday"], hol)
        else:
            rh, rs = orc.esx_neighbourhood_deltas(a[0][0][:T], ids, D, S, kw["start_weekday"], hol, kw["skills"])
        assert np.array_equal(a[1], rh) and np.array_equal(a[2], rs), (D, S)
