"""CPU oracle for the event-replay path. TEST INFRASTRUCTURE ONLY — see oracle/sgr_oracle.h."""
