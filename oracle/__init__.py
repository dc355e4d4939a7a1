"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the reference VideoTGB hot path.

Nothing under ``videotgb_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker or the timed CPU baseline.

Parity pinning: the reference ships no tests or golden vectors.
Every function here is therefore pinned against *outputs of the reference's own code
run in the build container*: ``oracle/gen_golden.py`` imports the reference classes
from ``/root/reference`` (through ``oracle/reference_shim.py``) or executes the
reference's inline selection statements read from the reference file at run time,
and writes the fixtures in ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
this package against those fixtures on CPU.
"""
