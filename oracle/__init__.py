"""CPU oracle for the pgeof hot path -- test infrastructure only (see ref_numpy.py)."""
