"""Throughput of the full CHANGE neighbourhood (rows[col] = v, n*(n-1) candidates per chain-step;
the reference's own move kind, examples/nqueens/src/lib.rs:227-229)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import constraint_solver_b200 as cs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
eng = cs.NQueensChains(n, chains, seed=42, neighbourhood=cs.CHANGE)
eng.init_random()
eng.step(1)
tot_m, tot_ms = 0, 0.0
for _ in range(3):
    st = eng.step(1)
    tot_m += st.moves_scored; tot_ms += st.device_ms
print(f"change neighbourhood n={n} chains={chains}: {tot_m / (tot_ms * 1e-3):.3e} moves/s ({tot_ms / 3:.2f} ms per step)")
