"""Time to score 0 for n-queens boards with the full swap neighbourhood (steepest descent with
sideways moves, LocalSearch::execute semantics) on `chains` restart chains at once.
usage: python scripts/nq_time_to_zero.py [n] [chains]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import constraint_solver_b200 as cs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 148
eng = cs.NQueensChains(n, chains, seed=42)
eng.init_random()
s0 = eng.scores()
t0 = time.perf_counter()
total_steps, moves = 0, 0
while True:
    st = eng.step(64)                      # 64 accepted moves per chain per launch
    total_steps += 64
    moves += st.moves_scored
    if st.chains_at_best or total_steps > 200_000:
        break
dt = time.perf_counter() - t0
print(f"n={n} chains={chains}: start score min/mean {int(s0.min())}/{float(s0.mean()):.0f}; "
      f"first chain at score 0 after {total_steps} steps, {dt:.3f} s, {moves:.3e} moves scored "
      f"({moves / dt:.3e} moves/s), chains at 0: {st.chains_at_best}")
