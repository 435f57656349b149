"""A SECOND, independently written scorer for the employee-scheduling score (test infrastructure).

Written directly from examples/employee-scheduling/src/lib.rs:148-222 and :261-375 in the
reference's own shape -- real calendar dates (datetime.date plays chrono::NaiveDate), a
dict-of-sets holiday table (HashMap<Employee, HashSet<Holiday>>), itertools-style windows and
Counter (itertools `counts()`), min/max over dict values -- and NOT from oracle/cs_oracle.c,
which works on day indices and weekday arithmetic.  The reference has no scheduling test, so
the C oracle's scheduling half cannot be pinned by a reference artefact; this rewrite at least
cross-checks it on thousands of random rotas (tests/test_oracle_cpu.py).
"""
from __future__ import annotations

import datetime as dt
from collections import Counter, defaultdict

SAT, SUN = 5, 6  # datetime.date.weekday(): Monday = 0


class ScheduleSolution:
    """lib.rs:127-192: start_date, end_date, date_to_employee (one entry per day, possibly a
    phantom entry past end_date, lib.rs:405-412)."""

    def __init__(self, start_date: dt.date, end_date: dt.date, date_to_employee):
        self.start_date, self.end_date = start_date, end_date
        self.date_to_employee = list(date_to_employee)

    def get_date_index(self, date):                       # :148-155
        if date < self.start_date or date > self.end_date:
            return None
        return (date - self.start_date).days

    def get_employee_for_date(self, date):                # :164-167
        index = self.get_date_index(date)
        return None if index is None else self.date_to_employee[index]

    def get_days_to_employees(self):                      # :181-191
        result = []
        current = self.start_date
        index = 0
        while True:
            result.append((current, self.date_to_employee[index]))
            if current >= self.end_date:
                break
            current += dt.timedelta(days=1)
            index += 1
        return result

    def get_employees_to_days(self):                      # :169-179
        result = defaultdict(list)
        for date, employee in self.get_days_to_employees():
            result[employee].append(date)
        return dict(result)


def is_weekend(date):                                     # :220-222
    return date.weekday() in (SAT, SUN)


def windows(seq, size):
    """slice::windows: every contiguous run of `size` items; none when the slice is shorter."""
    return [seq[i:i + size] for i in range(len(seq) - size + 1)]


def get_weekday_to_employee_counts_score(solution):       # :194-218
    day_counts = {}
    for date, employee in solution.get_days_to_employees():
        if is_weekend(date):
            continue
        day_counts.setdefault(date.weekday(), Counter())[employee] += 1
    score = 0
    for _day, employee_count in day_counts.items():
        if len(employee_count) <= 1:
            continue
        score += min(employee_count.values())             # MinMax(min, _max) => += min
    return score


def get_scored_solution(solution, employee_to_holidays):  # :261-375
    """employee_to_holidays: {employee: set of datetime.date}.  Raises KeyError-like ValueError
    where the reference unwrap()s a None (:275)."""
    hard = soft = 0
    for employee, holidays in employee_to_holidays.items():           # :273-280
        for holiday in holidays:
            actual = solution.get_employee_for_date(holiday)
            if actual is None:
                raise ValueError("holiday outside the schedule (reference panics)")
            if actual == employee:
                hard += 1
    days_to_employees = solution.get_days_to_employees()
    employees_to_days = solution.get_employees_to_days()
    for w in windows(days_to_employees, 2):                           # :286-292
        if w[0][1] == w[1][1]:
            hard += 1
    for w in windows(days_to_employees, 9):                           # :295-315
        d1, d2, d3, d4 = w[0], w[1], w[7], w[8]
        if not (is_weekend(d1[0]) and is_weekend(d2[0])):
            continue
        hard += (d1[1] == d3[1]) + (d1[1] == d4[1]) + (d2[1] == d3[1]) + (d2[1] == d4[1])
    for w in windows(days_to_employees, 14):                          # :318-327
        hard += sum(1 for c in Counter(e for _, e in w).values() if c > 3)
    for w in windows(days_to_employees, 7):                           # :330-339
        soft += sum(1 for c in Counter(e for _, e in w).values() if c > 2)
    soft += get_weekday_to_employee_counts_score(solution)            # :342
    lens = [len(days) for days in employees_to_days.values()]         # :345-351
    if len(lens) >= 2:                                                # MinMaxResult::MinMax needs 2 items
        soft += max(lens) - min(lens)
    wk = [sum(1 for d in days if is_weekend(d)) for days in employees_to_days.values()]  # :354-365
    if len(wk) >= 2:
        soft += max(wk) - min(wk)
    return hard, soft


# ---------------------------------------------------------------- the slot-generalised EXTENSION
def get_scored_solution_slots(start_date, n_days, shifts_per_day, slot_to_employee, employee_to_holidays,
                              employee_to_skills=None):
    """Independent rewrite of the EXTENSION's definition (oracle/cs_oracle.c: esx_terms; not pinned
    by the reference): the rota is a list of (date, shift, employee) triples, day-major.  Returns
    (hard, soft, terms[10]).  employee_to_skills: {employee: set of shift kinds}; None = all."""
    S = shifts_per_day
    rota = [(start_date + dt.timedelta(days=t // S), t % S, slot_to_employee[t]) for t in range(n_days * S)]
    days = [start_date + dt.timedelta(days=d) for d in range(n_days)]
    by_day = {d: [e for (dd, _s, e) in rota if dd == d] for d in days}
    terms = [0] * 10
    for employee, holidays in employee_to_holidays.items():              # H1: every slot of a holiday
        for holiday in holidays:
            if holiday not in by_day:
                raise ValueError("holiday outside the schedule")
            terms[0] += sum(1 for e in by_day[holiday] if e == employee)
    for w in windows(rota, 2):                                            # H2: consecutive slots
        terms[1] += w[0][2] == w[1][2]
    for w in windows(days, 9):                                            # H3: per shift kind
        if not (is_weekend(w[0]) and is_weekend(w[1])):
            continue
        for s in range(S):
            a, b, c, d = by_day[w[0]][s], by_day[w[1]][s], by_day[w[7]][s], by_day[w[8]][s]
            terms[2] += (a == c) + (a == d) + (b == c) + (b == d)
    for w in windows(days, 14):                                           # H4: > 3 slots in 14 days
        terms[3] += sum(1 for c in Counter(e for d in w for e in by_day[d]).values() if c > 3)
    for w in windows(days, 7):                                            # S1: > 2 slots in 7 days
        terms[4] += sum(1 for c in Counter(e for d in w for e in by_day[d]).values() if c > 2)
    per_weekday = {}
    for d, _s, e in rota:                                                 # S2
        if not is_weekend(d):
            per_weekday.setdefault(d.weekday(), Counter())[e] += 1
    for counts in per_weekday.values():
        if len(counts) > 1:
            terms[5] += min(counts.values())
    totals = Counter(e for _d, _s, e in rota)                             # S3 / S4
    weekend = {e: sum(1 for d, _s, x in rota if x == e and is_weekend(d)) for e in totals}
    if len(totals) >= 2:
        terms[6] = max(totals.values()) - min(totals.values())
        terms[7] = max(weekend.values()) - min(weekend.values())
    for d in days:                                                        # X1: same-day overlap pairs
        terms[8] += sum(c * (c - 1) // 2 for c in Counter(by_day[d]).values())
    if employee_to_skills is not None:                                    # X2: skill
        terms[9] = sum(1 for _d, s, e in rota if s not in employee_to_skills.get(e, set()))
    return sum(terms[:4]) + terms[8] + terms[9], sum(terms[4:8]), terms
