"""TEST INFRASTRUCTURE ONLY.

CPU restatements (torch fp32 / numpy float64) of the reference algorithms on the
DiariZen inference hot path.  Nothing in `diarizen_b200/` imports this package:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl
reference legs may.  Each function cites the reference file:line it follows.
"""
