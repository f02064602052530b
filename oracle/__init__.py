"""CPU oracle (test infrastructure only; see pq_oracle.py)."""
