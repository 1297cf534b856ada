"""CPU oracle (test infrastructure only -- see saeno_oracle.py / saeno_oracle.c headers)."""
