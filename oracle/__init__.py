"""CPU oracle of the racing-environment step -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package, and only as the checker or the CPU baseline.  The product
(``racing_dreamer_b200``) never imports it.

Contents
* ``rd_oracle.c``   float64/integer C restatement of every stage (the specification for the [NEW-SPEC] stages).
* ``np_oracle.py``  float64 NumPy rendition of the dynamics and the LiDAR traversal (cross-check of the C file).
* ``ref_stubs.py``  sys.modules stubs that let the UNMODIFIED reference ``dreamer/wrappers.py`` import in the
                    build container (used by tests/golden/make_golden.py; /root/reference does not travel).

Parity status: a3/a4/a5/a9/a10/a11 pinned against the reference's own classes via tests/golden;
a1/a2/a7/a8 **parity unpinned** (arithmetic lives in un-vendored racecar_gym + pybullet) -- see rd_oracle.c.
"""
from .binding import Oracle, OracleMap, build_oracle, default_config  # noqa: F401
