"""CPU oracle for wsi_b200 -- TEST INFRASTRUCTURE ONLY (see oracle/wsi_oracle.c)."""
