"""CPU oracle (test infrastructure).  PARITY UNPINNED: see oracle/xpbd_oracle.c."""
