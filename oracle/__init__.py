"""CPU checker (test infrastructure only) — see oracle/oracle.py."""
