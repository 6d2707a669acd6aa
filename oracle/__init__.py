"""CPU oracle package (test infrastructure only -- see oracle/oco_oracle.c)."""
