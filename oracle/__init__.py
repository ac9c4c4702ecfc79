"""TEST INFRASTRUCTURE -- CPU oracle package (checker only; see oracle/sc_oracle.h)."""
