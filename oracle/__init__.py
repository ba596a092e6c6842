"""CPU oracle for the DR4SR training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dr4sr_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and there only as the checker / the CPU arm, never as the product path.
"""
