"""CPU oracle for the VINS-RGBD-FAST hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import or execute it, and only as the checker
(or as the timed CPU reference arm) -- never on the product path.
"""
