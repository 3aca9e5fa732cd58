"""CPU oracle for the rollout hot path -- TEST INFRASTRUCTURE ONLY (see msw_oracle.h)."""
