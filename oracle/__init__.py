"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/tealeaf_oracle.c)."""
