"""CPU oracle (test infrastructure only) -- see sleekit_oracle.py."""
