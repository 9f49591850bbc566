"""TEST INFRASTRUCTURE — host oracle for the SpGEMM hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``pem_spgemm_b200``) never does and has no CPU fallback.

* ``oracle.host``  : OpenMP row-wise Gustavson (C++ in ``gustavson.cpp``, built by
  ``make -C oracle``) behind ctypes.
* ``oracle.tiles`` : numpy restatement of the reference's tiled-CSR data contract and
  of its step-1/step-2 intermediate arrays.

Parity pins: the reference has no tests or golden vectors of its own; the oracle is
pinned by (a) config-1 known answers, (b) scipy.sparse as an
independent CPU implementation, and (c) dumps of the reference's own sm_100 rebuild
(``oracle/_ref``) run on a B200 and committed under ``tests/golden/``.
"""
