"""CPU oracle for the probabilit sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU arm.
``probabilit_b200`` never imports this package (tests/test_abi_cpu.py
enforces that).

Each function is a NumPy/SciPy restatement of the reference algorithm and cites
the reference ``file:line`` it follows (paths relative to /root/reference).

Parity pinning: the restatements are pinned against (a) the reference's own doctest
/ README golden values (tests/test_oracle_golden.py) and (b) golden vectors produced
by importing the *unmodified* reference in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz).
"""
