"""CPU oracle for the Relation-DETR hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package, and only as the checker.  The product package
(``relation-detr_b200/``) never imports it.  See ``rdetr_oracle.c`` for the parity status
(pinned against fixtures generated from the reference by ``make_golden.py``).
"""
