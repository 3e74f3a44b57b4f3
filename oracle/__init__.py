"""CPU oracle for the RCPS hot path - TEST INFRASTRUCTURE ONLY (see oracle/rcps_oracle.py)."""
