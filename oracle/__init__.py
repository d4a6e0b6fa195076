"""CPU oracle for the aslsd LSE hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy restatement of the reference algorithm
(saadlabyad/aslsd) for the path named in BASELINE.json: kappa/varpi index
maps, the exact LSE sums, and the stratified (ASLSD) per-stratum statistics.
Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or the CPU arm being timed).  Nothing under ``aslsd_b200/`` imports
it: the product path has no CPU fallback.

Parity pinning: the reference ships no golden vectors or runnable tests for
this path, so the oracle is pinned against outputs
of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports /root/reference) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the oracle
against them, and ``tests/test_oracle_vs_reference.py`` re-runs the live
comparison whenever /root/reference is present.
"""
