"""CPU oracle package marker (test infrastructure only; see oracle/oracle_math.h)."""
