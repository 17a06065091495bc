"""CPU oracle of the DCCN hot path -- test infrastructure only (see dccn_oracle.py)."""
