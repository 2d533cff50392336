"""CPU oracle for the batched RJ-MCMC hot path of BayesBay 0.3.11.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a plain numpy/scipy restatement of the reference algorithm (every function cites
the reference ``file:line`` it follows, paths relative to ``/root/reference``).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it -- and only as the *checker* or as the timed CPU
baseline, never as something the CUDA product path falls back to.  Nothing under
``bayesbay_b200/`` imports ``oracle``.

Parity status: PINNED.  The oracle is checked (``tests/test_oracle_golden.py``) against
fixtures under ``tests/golden/`` that were produced by running the *unmodified
reference* in the build container (``tests/golden/make_golden.py``, which imports
``bayesbay`` from a build of ``/root/reference`` and feeds it a recorded random tape),
plus the reference's own known-answer vectors (tests/20_test_param_space_log_prior.py:16-30,
the ``Voronoi1D.compute_cell_extents`` docstring at discretization/_voronoi.py:657-666,
the quickstart prior values of tests/22_log_like_func.py:24-31).

Third-party arithmetic on the path that is not under /root/reference (restated here):
``scipy.spatial.KDTree.query`` (scipy>=1.9, here 1.18.1) == exact Euclidean nearest
neighbour -> brute-force ``argmin(dx*dx+dy*dy)`` with lowest-index tie-break;
``scipy.sparse`` CSR mat-vec; CPython ``random`` and legacy ``numpy.random`` draws
(restated as counter/tape driven draws, see :mod:`oracle.rng`).
"""
from . import rng, kernels, chain  # noqa: F401
