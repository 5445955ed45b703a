"""CPU oracle (test infrastructure only) — see oracle/nlsh_oracle.py."""
