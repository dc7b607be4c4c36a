"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/nmf_oracle.c header)."""
