"""CPU oracle for the MA-CJD hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (NumPy float64 for the
environment physics, plain eager PyTorch fp32/fp64 for the networks) of the
algorithm of the reference repository, written from its behaviour and pinned
against outputs of the *unmodified* reference run in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and only as
the checker / the CPU baseline.  The product package never imports it; the
product path fails loudly if the CUDA library is missing.
"""
