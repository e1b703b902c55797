"""CPU oracle for the hot path (test infrastructure; see ref_port.py / gmr_oracle.c headers)."""
