"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/cs_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
