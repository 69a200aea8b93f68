"""CPU oracle for the Exposure hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``exposure_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference (yuanming-hu/exposure @ 7bb838a) ships no golden
vectors / known-answer tests, and its TensorFlow-1.6 graph cannot be executed in
this image (no tensorflow; util.py:658 does not parse on Python >= 3.7).  The
oracle is therefore a formula-level restatement of the reference sources, each
function citing the file:line it follows, self-checked by fp32-vs-fp64 agreement,
finite differences and algebraic invariants (tests/test_oracle_*.py).
"""
