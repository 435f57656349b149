#!/usr/bin/env python
"""Driver with the reference CLI's flags (examples/nqueens/src/main.rs:95-150):
  --seed/-s STRING (default "42"), --board-size/-b INT (default 8)
and its solver constants (:129-135), running on the B200 evaluator: `--chains` ILS chains at
once, stop when one reaches score 0.  The seed string is hashed like hash_str (:28-33,
blake2b-256) and the first 8 bytes key the Philox streams (the reference keys ChaCha20)."""
import argparse
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import constraint_solver_b200 as cs  # noqa: E402


def hash_str(seed: str) -> int:
    return int.from_bytes(hashlib.blake2b(seed.encode(), digest_size=32).digest()[:8], "little")


def main():
    ap = argparse.ArgumentParser(description="Local Search N-Queens Example (B200)")
    ap.add_argument("-s", "--seed", default="42", help="Random seed, any string")
    ap.add_argument("-b", "--board-size", type=int, default=8, help="Board size")
    ap.add_argument("--chains", type=int, default=1024, help="independent ILS chains on the GPU")
    args = ap.parse_args()
    print("local search n-queens example")
    n = args.board_size
    eng = cs.NQueensChains(n, args.chains, seed=hash_str(args.seed), neighbourhood=cs.CHANGE)
    eng.init_random()
    eng.ils_init(32)                                   # best_solutions_capacity
    st = eng.ils_run(10_000, 10_000, 5, stop_when_any_best=True)  # ILS rounds, LS iterations, no-improve
    rows, score = eng.ils_best(st["best_chain"])
    if n <= 32:
        print("result.solution:\n%r" % cs.NQueensSolution(rows))
    else:
        print("result.solution rows:", rows.tolist())
    print("result.score: NQueensScore(%d)" % score)
    print("rounds: %d  chains solved: %d  moves scored: %d" % (st["rounds_run"], st["chains_done"], st["moves_scored"]))


if __name__ == "__main__":
    main()
