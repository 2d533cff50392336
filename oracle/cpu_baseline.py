"""CPU baseline: the oracle's chain step run the way the reference runs its chains.

TEST / BENCH INFRASTRUCTURE (see ``oracle/__init__.py``) -- used only by ``bench.py`` (the
``cpu_baseline`` leg and ``--impl reference``).

The reference advances ``n_chains`` independent ``MarkovChain`` objects, one OS process per chain, with
``joblib.Parallel(n_jobs=len(chains), backend="loky")`` (src/bayesbay/samplers/_samplers.py:232-235).
The reference is a Python package and cannot travel to the GPU box, so its path is timed through this
port (``kind: "port"``): the same per-step work -- Python-level proposal, a KD-tree query of every grid
point (``scipy.spatial.KDTree``, what the tomography forward of the tutorials calls), a scipy CSR
mat-vec, the Gaussian misfit -- on one process per host core.  In the survey container the unmodified
reference ran the C3-shaped problem at ~200 chain-steps/s per core; this port
measures the same order.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import chain as OC
from . import kernels as OK
from . import rng as OR


def _install_reference_style_forward(spec):
    """Use the reference's own numerical routines for the forward (KD-tree + scipy CSR) instead of the
    brute-force restatement (identical results, tests/test_oracle_golden.py; much faster on the CPU)."""
    import scipy.sparse
    import scipy.spatial

    jac = {}
    for i, t in enumerate(spec["targets"]):
        f = t["forward"]
        if f["kind"] == "ray2d":
            jac[i] = scipy.sparse.csr_matrix((f["data"], f["indices"], f["indptr"]),
                                             shape=(f["indptr"].size - 1, f["points"].shape[0]))

    def forward_target(t, st):
        f = t["forward"]
        sp = st.spaces[f["space"]]
        if f["kind"] == "ray2d":
            i = [j for j, tt in enumerate(spec["targets"]) if tt is t][0]
            nearest = scipy.spatial.KDTree(sp.sites).query(f["points"])[1]
            field = sp.values[f["param"]][nearest]
            return jac[i] @ (1.0 / field if f.get("transform", "reciprocal") == "reciprocal" else field)
        if f["kind"] == "lookup1d":
            return OK.lookup1d_forward(f["x"], sp.sites, sp.values[f["param"]])
        if f["kind"] == "polynomial":
            m = sp.values[f["param"]]
            return np.vander(f["x"], len(m), True) @ m
        return OK.linear_forward(f["G"], np.concatenate([sp.values[p] for p in f["params"]]))

    OC.forward_target = forward_target


def _worker(args):
    spec, n_steps, seed, warm = args
    _install_reference_style_forward(spec)
    gen = np.random.default_rng(seed)
    s = OR.NumpyStream(gen)
    st = OC.init_state(spec, s)
    for _ in range(warm):
        st, _ = OC.step(spec, st, s)
    t0 = time.perf_counter()
    acc = 0
    for _ in range(n_steps):
        st, info = OC.step(spec, st, s)
        acc += info["accepted"]
    return time.perf_counter() - t0, acc


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def run_pool(spec, n_steps: int, n_procs: int = None, warm: int = 5, seed: int = 0):
    """``n_procs`` chains (one process each) x ``n_steps`` iterations.  Returns
    ``(chain_steps_per_s, n_procs, wall_seconds)``: aggregate rate = n_procs * n_steps / wall of the timed
    call (process start-up and the warm-up iterations excluded, like the reference protocol of BASELINE.md
    section 4)."""
    import multiprocessing as mp

    n_procs = n_procs or host_cores()
    if n_procs == 1:
        dt, _ = _worker((spec, n_steps, seed, warm))
        return n_steps / dt, 1, dt
    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs) as pool:
        t0 = time.perf_counter()
        out = pool.map(_worker, [(spec, n_steps, seed + i, warm) for i in range(n_procs)])
        wall = time.perf_counter() - t0
    slowest = max(dt for dt, _ in out)
    return n_procs * n_steps / slowest, n_procs, wall


def steps_for_budget(spec, seconds: float, seed: int = 0):
    """Pick a step count that makes one process run for about ``seconds``."""
    dt, _ = _worker((spec, 5, seed, 2))
    per_step = max(dt / 5, 1e-6)
    return max(5, int(seconds / per_step))
