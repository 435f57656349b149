#!/usr/bin/env python
"""Driver mirroring examples/employee-scheduling/src/main.rs (7 employees, 2022-05-09 + 30 days,
no holidays, its solver constants :24-31) and, with --json, the wasm API's JSON shapes
(web/employee-scheduling-wasm-bindgen/src/lib.rs:86-110): input
{"startDate","endDate","employees":[{"id":..}],"employeeHolidays":[["YYYY-MM-DD",..],..]} and output
{"score":{"hard_score","soft_score"},"days_to_employees":[["Mon 2022-05-09",{"id":..}],..]}."""
import argparse
import datetime as dt
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import constraint_solver_b200 as cs  # noqa: E402


def hash_str(seed: str) -> int:
    return int.from_bytes(hashlib.blake2b(seed.encode(), digest_size=32).digest()[:8], "little")


def solve(inp, chains=256, seed="42"):
    start = dt.date.fromisoformat(inp["startDate"])
    end = dt.date.fromisoformat(inp["endDate"])
    D = (end - start).days + 1
    ids = [e["id"] for e in inp["employees"]]
    hol = [(e, (dt.date.fromisoformat(h) - start).days)
           for e, hs in zip(ids, inp.get("employeeHolidays", [[] for _ in ids])) for h in hs]
    eng = cs.ScheduleChains(D, ids, start_weekday=start.weekday(), holidays=hol, n_chains=chains,
                            seed=hash_str(seed))
    eng.init_random()
    eng.ils_init(64)                                    # best_solutions_capacity
    st = eng.ils_run(250, 1_000, 20, stop_when_any_best=True)
    rows, hard, soft = eng.ils_best(st["best_chain"])
    days = [((start + dt.timedelta(days=i)).strftime("%a %Y-%m-%d"), {"id": int(rows[i])}) for i in range(D)]
    return {"score": {"hard_score": float(hard), "soft_score": float(soft)}, "days_to_employees": days}, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", help="EmployeeSchedulingInput JSON file (default: the reference's hard-coded instance)")
    ap.add_argument("--chains", type=int, default=256)
    args = ap.parse_args()
    print("employee scheduling local search example")
    if args.json:
        inp = json.load(open(args.json))
    else:
        inp = {"startDate": "2022-05-09", "endDate": "2022-06-08",
               "employees": [{"id": i} for i in range(7)], "employeeHolidays": [[] for _ in range(7)]}
    out, st = solve(inp, args.chains)
    print(json.dumps(out, indent=1))
    print("rounds: %d  chains at (0,0): %d  moves scored: %d" % (st["rounds_run"], st["chains_done"], st["moves_scored"]))


if __name__ == "__main__":
    main()
