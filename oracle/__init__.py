"""oracle/ - TEST INFRASTRUCTURE ONLY.

A CPU (numpy) restatement of the reference's Monte Carlo algorithm, used solely as the
checker for the CUDA path:
  * tests/ compare the kernels with it on the same seeded inputs,
  * __graft_entry__.smoke() checks one small run against it,
  * bench.py's cpu_baseline / --impl reference legs time it.
Nothing under montecarlo-risk-engine_b200/ imports it; the product path fails loudly
without the CUDA library instead of falling back to this code.

Parity pinning: the reference is pure Python and runs in the build container, so the
oracle is pinned against outputs of the reference itself (tests/golden/*.json, generated
by tests/golden/make_golden.py which imports /root/reference/src) including the
reference's own seeded known-answer values (tests/pytests/test_cva.py:188-189).
Each function cites the reference file:line it restates.
"""
