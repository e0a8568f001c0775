"""CPU oracle for the GARLIC hot path — TEST INFRASTRUCTURE ONLY (see garlic_oracle.c)."""
