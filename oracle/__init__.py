"""CPU oracle (test infrastructure only): see nde_oracle.py / nde_oracle.c headers. PARITY UNPINNED."""
