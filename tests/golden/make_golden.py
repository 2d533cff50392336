#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

What it does
------------
1. Copies /root/reference to a scratch directory, builds its Cython helper with the
   reference's own setup.py, and puts on-disk stub packages for ``matplotlib`` and
   ``shapely`` (hard imports at src/bayesbay/discretization/_voronoi.py:7-11, used only for
   plotting / polygon domains) on the path.  Nothing from the reference is copied into
   this repository; only numeric outputs are stored.
2. Monkey-patches ``random.*`` and ``numpy.random.*`` with tape readers that follow the
   rules documented in ``oracle/rng.py`` -- the reference then consumes a recorded
   stream of uniforms, which makes its chains replayable by the oracle and by the CUDA
   kernels (tape mode).
3. Records, for small instances of the BASELINE.json configurations,
   * deterministic kernel outputs on random states (``kernels.npz``),
   * full step replays: tape, per-step log_prob_ratio / log_like_ratio / accept / state
     (``replay_<config>.npz``),
   * parallel-tempering swap rounds (``pt_swap.npz``),
   * long-run statistics (acceptance per move type, k histogram, sigma mean)
     (``stats_<config>.npz``) for Monte-Carlo-level checks.
"""
from __future__ import annotations

import math
import os
import random
import shutil
import subprocess
import sys
from bisect import bisect_right

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
SCRATCH = os.environ.get("BBB_REF_SCRATCH", "/tmp/bbb_ref_build")
REFERENCE = "/root/reference"


# --------------------------------------------------------------------------- #
# 1. build + import the reference
# --------------------------------------------------------------------------- #
def import_reference():
    src = os.path.join(SCRATCH, "src")
    if not os.path.isdir(src) or not any(f.endswith(".so") for f in os.listdir(os.path.join(src, "bayesbay"))):
        if os.path.isdir(SCRATCH):
            shutil.rmtree(SCRATCH)
        os.makedirs(SCRATCH)
        for name in ("src", "setup.py", "README.md", "pyproject.toml", "MANIFEST.in"):
            p = os.path.join(REFERENCE, name)
            (shutil.copytree if os.path.isdir(p) else shutil.copy)(p, os.path.join(SCRATCH, name))
        subprocess.run("chmod -R u+w .", shell=True, cwd=SCRATCH, check=True)
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=SCRATCH, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    stubs = os.path.join(SCRATCH, "stubs")
    os.makedirs(os.path.join(stubs, "matplotlib"), exist_ok=True)
    os.makedirs(os.path.join(stubs, "shapely"), exist_ok=True)
    for rel, body in (("matplotlib/__init__.py", ""), ("matplotlib/pyplot.py", ""), ("shapely/__init__.py", ""),
                      ("shapely/geometry.py", "class Polygon:\n    pass\nclass Point:\n    pass\n")):
        with open(os.path.join(stubs, rel), "w") as f:
            f.write(body)
    sys.path.insert(0, stubs)
    sys.path.insert(0, src)
    import bayesbay  # noqa

    return bayesbay


# --------------------------------------------------------------------------- #
# 2. tape-driven random draws (rules of oracle/rng.py)
# --------------------------------------------------------------------------- #
class Tape:
    def __init__(self, seed, n=1 << 16):
        self.values = np.random.Generator(np.random.PCG64(seed)).random(n)
        self.n = 0

    def u(self):
        v = float(self.values[self.n])
        self.n += 1
        return v


TAPE = None  # the active tape


def _std_normal():
    u1 = TAPE.u()
    u2 = TAPE.u()
    return math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(2.0 * math.pi * u2)


def _p_random():
    return 1.0 - TAPE.u()


def _p_uniform(a, b):
    return a + (b - a) * TAPE.u()


def _p_randint(a, b):
    if a == b:
        return a
    n = b - a + 1
    return a + min(int(TAPE.u() * n), n - 1)


def _p_choices(population, weights=None, *, cum_weights=None, k=1):
    assert k == 1 and weights is not None
    population = list(population)
    if len(population) == 1:
        return [population[0]]
    cum = list(np.cumsum(np.asarray(weights, dtype=float)))
    return [population[min(bisect_right(cum, TAPE.u() * cum[-1]), len(population) - 1)]]


def _p_normalvariate(mu=0.0, sigma=1.0):
    return mu + sigma * _std_normal()


def _np_uniform(low=0.0, high=1.0, size=None):
    if size is None:
        return low + (high - low) * TAPE.u()
    n = int(np.prod(size))
    low_b = np.broadcast_to(np.asarray(low, dtype=float), (n,))
    high_b = np.broadcast_to(np.asarray(high, dtype=float), (n,))
    return np.array([low_b[i] + (high_b[i] - low_b[i]) * TAPE.u() for i in range(n)]).reshape(size)


def _np_normal(loc=0.0, scale=1.0, size=None):
    if size is None:
        return loc + scale * _std_normal()
    n = int(np.prod(size))
    loc_b = np.broadcast_to(np.asarray(loc, dtype=float), (n,))
    scale_b = np.broadcast_to(np.asarray(scale, dtype=float), (n,))
    return np.array([loc_b[i] + scale_b[i] * _std_normal() for i in range(n)]).reshape(size)


def _np_laplace(loc=0.0, scale=1.0, size=None):
    """numpy's inverse-CDF rule on one uniform per value (oracle/rng.py draw_laplace)."""
    def one(mu, b):
        u = TAPE.u()
        if u >= 0.5:
            return mu - b * math.log(2.0 - u - u)
        if u > 0.0:
            return mu + b * math.log(u + u)
        return mu

    if size is None:
        return one(loc, scale)
    n = int(np.prod(size))
    loc_b = np.broadcast_to(np.asarray(loc, dtype=float), (n,))
    scale_b = np.broadcast_to(np.asarray(scale, dtype=float), (n,))
    return np.array([one(loc_b[i], scale_b[i]) for i in range(n)]).reshape(size)


def _np_choice(a, size=None, replace=True, p=None):
    assert size == 2 and replace is False
    n = len(a)
    i = min(int(TAPE.u() * n), n - 1)
    j = min(int(TAPE.u() * (n - 1)), n - 2)
    if j >= i:
        j += 1
    return [a[i], a[j]]


def install_patches():
    random.random = _p_random
    random.uniform = _p_uniform
    random.randint = _p_randint
    random.choices = _p_choices
    random.normalvariate = _p_normalvariate
    random.gauss = _p_normalvariate
    np.random.uniform = _np_uniform
    np.random.normal = _np_normal
    np.random.laplace = _np_laplace
    np.random.choice = _np_choice


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def reference_forward(ing):
    """Plain reference forward callables, written as the tutorials write them."""
    import scipy.sparse
    from bayesbay.discretization import Voronoi1D

    kind = ing["kind"]
    if kind == "tomography_2d":
        jac = scipy.sparse.csr_matrix((ing["data"], ing["indices"], ing["indptr"]),
                                      shape=(ing["indptr"].size - 1, ing["points"].shape[0]))
        points = ing["points"]

        def forward(state):  # 31_sw_tomography.ipynb cell 12
            voronoi = state["voronoi"]
            kdtree = voronoi.load_from_cache("kdtree")
            nearest = kdtree.query(points)[1]
            interp_vel = voronoi.get_param_values("vel")[nearest]
            state.save_to_extra_storage("interp_vel", interp_vel)  # as the notebook does (cell 12)
            return jac @ (1 / interp_vel)

        return forward
    if kind == "partition_1d":
        x = ing["x"]

        def forward(state):  # library nearest-nuclei lookup, discretization/_voronoi.py:705-732
            voro = state["voronoi"]
            return Voronoi1D.interpolate_tessellation(voro["discretization"], voro["y"], x, "nuclei")

        return forward
    if kind == "transd_polyfit":
        x = ing["x"]

        def forward(state):  # 03_transd_polyfit.ipynb cell 5
            m = state["my_param_space"]["coefficients"]
            return np.vander(x, len(m), True) @ m

        return forward
    if kind == "polyfit":
        G = ing["G"]

        def forward(state):  # 02_hierarchical_polyfit.ipynb cell 12
            m = [state["my_param_space"][f"m{i}"] for i in range(4)]
            return np.squeeze(G @ m)

        return forward
    raise ValueError(kind)


class Recorder:
    """Wraps a reference perturbation callable and records what it returned."""

    def __init__(self, func, log):
        self.func = func
        self.log = log

    @property
    def __name__(self):
        return self.func.__name__

    def __call__(self, state):
        new_state, ratio = self.func(state)
        self.log["lpr"] = ratio
        self.log["outer"] = self.func.__name__
        if self.func.__name__.startswith("ParamSpacePerturbation"):
            self.log["inner"] = next(iter(new_state.load_from_cache("perturb_stats")))
        else:
            self.log["inner"] = None
        return new_state, ratio


def snapshot_corr(state, target_name):
    if state[target_name] is None:
        return np.nan
    c = state[target_name].correlation
    return np.nan if c is None else c


def snapshot(state, ps_name, param_names, target_name, kmax, ndim):
    ps = state[ps_name]
    noise = state[target_name]
    k = ps.n_dimensions
    sites = np.full((kmax, max(ndim, 1)), np.nan)
    if ndim > 0:
        sites[:k] = np.asarray(ps["discretization"]).reshape(k, ndim)
    vals = np.full((len(param_names), kmax), np.nan)
    for j, p in enumerate(param_names):
        vals[j, :k] = ps[p]
    return k, sites, vals, (np.nan if noise is None else noise.std)


def move_id(inner, outer):
    if outer.startswith("NoisePerturbation"):
        return 4
    if inner.startswith("BirthPerturbation"):
        return 0
    if inner.startswith("DeathPerturbation"):
        return 1
    if inner.endswith(".discretization)"):
        return 3
    return 2


def replay(bb, ing, n_chains, n_steps, seed0, name):
    """Tape-driven chains of the reference, recorded step by step."""
    global TAPE
    from bayesbay_b200 import synthetic

    ps_name = {"tomography_2d": "voronoi", "partition_1d": "voronoi", "polyfit": "my_param_space",
               "transd_polyfit": "my_param_space"}[ing["kind"]]
    ndim = {"tomography_2d": 2, "partition_1d": 1, "polyfit": 0, "transd_polyfit": 0}[ing["kind"]]
    out = None
    names = None
    for c in range(n_chains):
        TAPE = Tape(seed0 + c)
        parameterization, log_likelihood = synthetic.build_problem(ing, bb, reference_forward)
        ps = parameterization.parameter_spaces[ps_name]
        pnames = list(ps.parameters.keys())
        kmax = ps._n_dimensions_max
        tname = log_likelihood.targets[0].name
        inv = bb.BayesianInversion(parameterization, log_likelihood, n_chains=1)
        chain = inv.chains[0]
        log = {}
        chain.set_perturbation_funcs([Recorder(f, log) for f in chain.perturbation_funcs],
                                     chain.perturbation_weights)
        names = [f.__name__ for f in chain.perturbation_funcs]
        orig_llr = chain._log_likelihood_ratio

        def rec_llr(new_state, _orig=orig_llr, _log=log):
            v = _orig(new_state)
            _log["llr"] = v
            return v

        chain._log_likelihood_ratio = rec_llr
        if out is None:
            out = dict(
                k=np.zeros((n_chains, n_steps + 1), np.int32),
                sites=np.zeros((n_chains, n_steps + 1, kmax, max(ndim, 1))),
                values=np.zeros((n_chains, n_steps + 1, len(pnames), kmax)),
                sigma=np.zeros((n_chains, n_steps + 1)), corr=np.zeros((n_chains, n_steps + 1)),
                lpr=np.zeros((n_chains, n_steps)), llr=np.zeros((n_chains, n_steps)),
                accepted=np.zeros((n_chains, n_steps), np.int8), move=np.zeros((n_chains, n_steps), np.int8),
                cursor=np.zeros((n_chains, n_steps + 1), np.int64),
                tape=np.zeros((n_chains, 0)),
            )
            tapes = []
        k, s, v, sg = snapshot(chain.current_state, ps_name, pnames, tname, kmax, ndim)
        out["k"][c, 0], out["sites"][c, 0], out["values"][c, 0], out["sigma"][c, 0] = k, s, v, sg
        out["cursor"][c, 0] = TAPE.n
        out["corr"][c, 0] = snapshot_corr(chain.current_state, tname)
        chain.save_current_iteration = False
        for t in range(n_steps):
            log.clear()
            n_acc = chain.statistics["n_accepted_models_total"]
            chain._next_iteration()
            out["lpr"][c, t] = log["lpr"]
            out["llr"][c, t] = log.get("llr", 0.0)
            out["accepted"][c, t] = chain.statistics["n_accepted_models_total"] - n_acc
            out["move"][c, t] = move_id(log["inner"] or "", log["outer"])
            k, s, v, sg = snapshot(chain.current_state, ps_name, pnames, tname, kmax, ndim)
            out["k"][c, t + 1], out["sites"][c, t + 1], out["values"][c, t + 1], out["sigma"][c, t + 1] = k, s, v, sg
            out["cursor"][c, t + 1] = TAPE.n
            out["corr"][c, t + 1] = snapshot_corr(chain.current_state, tname)
        tapes.append(TAPE.values[: TAPE.n].copy())
    L = max(t.size for t in tapes)
    tape = np.zeros((n_chains, L))
    for c, t in enumerate(tapes):
        tape[c, : t.size] = t
    out["tape"] = tape
    out["perturbation_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, f"replay_{name}.npz"), **out)
    acc = out["accepted"].mean()
    print(f"replay_{name}: {n_chains} chains x {n_steps} steps, acceptance {acc:.3f}, tape len {L}, "
          f"moves {np.bincount(out['move'].ravel(), minlength=5)}")


# --------------------------------------------------------------------------- #
# 3. deterministic kernels
# --------------------------------------------------------------------------- #
def kernels_golden(bb):
    from bayesbay._utils_1d import (compute_voronoi1d_cell_extents, delete_1d, insert_1d,
                                    interpolate_nearest_1d, inverse_covariance, nearest_neighbour_1d)
    from bayesbay.discretization import Voronoi1D, Voronoi2D
    from bayesbay.likelihood import LogLikelihood, Target
    from bayesbay.parameterization import ParameterSpace
    from bayesbay.prior import GaussianPrior, LaplacePrior, UniformPrior
    from bayesbay import State, ParameterSpaceState, DataNoiseState
    from bayesbay_b200 import synthetic

    rng = np.random.Generator(np.random.PCG64(7))
    out = {}
    # --- F1 2-D raster: Voronoi2D.interpolate_tessellation (KD-tree) on random states
    pts = synthetic.regular_grid_points(23, 17)
    ks = [1, 2, 5, 37, 200]
    out["r2_points"] = pts
    out["r2_ks"] = np.array(ks)
    for i, k in enumerate(ks):
        sites = rng.uniform(-1, 1, (k, 2))
        vals = rng.uniform(2, 4, k)
        field = Voronoi2D.interpolate_tessellation(sites, vals, pts)
        import scipy.spatial

        idx = scipy.spatial.KDTree(sites).query(pts)[1]
        assert np.array_equal(field, vals[idx])
        out[f"r2_sites_{i}"], out[f"r2_vals_{i}"], out[f"r2_idx_{i}"] = sites, vals, idx.astype(np.int32)
    # --- F1/F4 1-D: interpolate_nearest_1d, nearest_neighbour_1d
    xq = np.concatenate((np.linspace(0, 10, 100), [-1.0, 0.0, 10.0, 11.0]))
    out["r1_x"] = xq
    ks1 = [1, 2, 9, 40]
    out["r1_ks"] = np.array(ks1)
    for i, k in enumerate(ks1):
        sites = np.sort(rng.uniform(0, 10, k))
        vals = rng.uniform(-35, 35, k)
        out[f"r1_sites_{i}"], out[f"r1_vals_{i}"] = sites, vals
        out[f"r1_field_{i}"] = np.asarray(interpolate_nearest_1d(xq, sites, vals))
        out[f"r1_nn_{i}"] = np.array([nearest_neighbour_1d(float(x), sites, int(sites.size)) for x in xq], np.int32)
    # exact ties / boundaries on hand-made sites
    sites = np.array([2.0, 5.5, 8.0, 10.0])
    xt = np.array([0.0, 2.0, 3.75, 5.5, 6.75, 8.0, 9.0, 10.0, 12.0])
    out["r1_tie_sites"], out["r1_tie_x"] = sites, xt
    out["r1_tie_nn"] = np.array([nearest_neighbour_1d(float(x), sites, 4) for x in xt], np.int32)
    out["r1_tie_field"] = np.asarray(interpolate_nearest_1d(xt, sites, np.arange(4.0)))
    # --- K: cell extents docstring examples (discretization/_voronoi.py:657-666)
    out["ext_a"] = Voronoi1D.compute_cell_extents(sites, lb=0, ub=None, fill_value=np.nan)
    out["ext_c"] = Voronoi1D.compute_cell_extents(sites, lb=0, ub=15, fill_value=np.nan)
    out["ins"] = np.asarray(insert_1d(np.arange(5.0), 2, 9.5))
    out["del"] = np.asarray(delete_1d(np.arange(5.0), 2))
    out["invcov"] = np.asarray(inverse_covariance(0.7, 0.3, 6))
    # --- L3 log prior: tests/20_test_param_space_log_prior.py:16-30 (custom prior term is 0 there)
    uni = UniformPrior("uniform", vmin=-1, vmax=1, perturb_std=0.1)
    gau = GaussianPrior("gaussian", mean=0, std=1, perturb_std=0.1)
    ps = ParameterSpace("ps", n_dimensions=1, parameters=[uni, gau])
    st = ParameterSpaceState(1, {"uniform": np.array([0.0]), "gaussian": np.array([0.0])})
    out["lp_test20"] = np.array(ps.log_prior(st))
    assert out["lp_test20"] == math.log(1 / 2) + math.log(1 / math.sqrt(2 * math.pi))
    lap = LaplacePrior("laplace", mean=0.5, scale=2.0, perturb_std=0.1)
    uni2 = UniformPrior("u", vmin=2, vmax=4, perturb_std=0.1)
    gau2 = GaussianPrior("g", mean=-3, std=2, perturb_std=0.1)
    ps2 = ParameterSpace("ps2", n_dimensions=7, parameters=[uni2, gau2, lap])
    vals = np.stack((rng.uniform(2, 4, 7), rng.normal(-3, 2, 7), rng.normal(0.5, 2, 7)))
    st2 = ParameterSpaceState(7, {"u": vals[0], "g": vals[1], "laplace": vals[2]})
    out["lp_vals"] = vals
    out["lp_joint"] = np.array(ps2.log_prior(st2))
    out["lp_quickstart"] = np.array([math.exp(UniformPrior("m", 5, 15, 0.4).log_prior(v)) for v in (5, 10, 20)])
    # prior ratios of value moves are exercised by the replays; record pdfs elementwise too
    out["lp_each"] = np.array([[p.log_prior(float(v)) for v in row] for p, row in zip((uni2, gau2, lap), vals)])
    # --- L1/L2 misfit, log-det, ratio
    n = 50
    dobs = rng.normal(0, 1, n)
    tgt = Target("t", dobs, std_min=0, std_max=1, std_perturb_std=0.1)
    dp_old = dobs + rng.normal(0, 0.2, n)
    dp_new = dobs + rng.normal(0, 0.2, n)
    ll = LogLikelihood(tgt, lambda s: None)
    ps_dummy = ParameterSpaceState(1, {"x": np.zeros(1)})
    s_old = State({"p": ps_dummy, "t": DataNoiseState(std=0.21, correlation=None)}, temperature=1)
    s_new = State({"p": ps_dummy.copy(), "t": DataNoiseState(std=0.19, correlation=None)}, temperature=1)
    s_old.save_to_cache("t.dpred", dp_old)
    s_new.save_to_cache("t.dpred", dp_new)
    out["ll_dobs"], out["ll_dp_old"], out["ll_dp_new"] = dobs, dp_old, dp_new
    out["ll_old"] = np.array(ll._get_misfit_and_det(s_old))
    out["ll_new"] = np.array(ll._get_misfit_and_det(s_new))
    out["ll_ratio_T1"] = np.array(ll.log_likelihood_ratio(s_old, s_new, 1.0))
    out["ll_ratio_T3"] = np.array(ll.log_likelihood_ratio(s_old, s_new, 3.0))
    # correlated hierarchical noise (tridiagonal inverse covariance)
    tgc = Target("t", dobs, std_min=0, std_max=1, std_perturb_std=0.1, noise_is_correlated=True)
    llc = LogLikelihood(tgc, lambda s: None)
    sc = State({"p": ps_dummy.copy(), "t": DataNoiseState(std=0.23, correlation=0.37)}, temperature=1)
    sc.save_to_cache("t.dpred", dp_old)
    out["ll_corr"] = np.array(llc._get_misfit_and_det(sc))
    # fixed (non-hierarchical) covariance: scalar and vector inverse covariance
    w = rng.uniform(0.5, 2.0, n)
    for tag, ci in (("s", 4.0), ("v", w)):
        tg = Target("t", dobs, covariance_mat_inv=ci)
        l2 = LogLikelihood(tg, lambda s: None)
        s0 = State({"p": ps_dummy}, temperature=1)
        s0.save_to_cache("t.dpred", dp_old)
        out[f"ll_fixed_{tag}"] = np.array(l2._get_misfit_and_det(s0))
    out["ll_w"] = w
    # --- F2 forward on the small tomography problem
    ing = synthetic.tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4, kmax=20)
    import scipy.sparse

    jac = scipy.sparse.csr_matrix((ing["data"], ing["indices"], ing["indptr"]), shape=(ing["indptr"].size - 1, 120))
    sites = rng.uniform(-1, 1, (11, 2))
    vel = rng.uniform(2, 4, 11)
    nearest = scipy.spatial.KDTree(sites).query(ing["points"])[1]
    out["fw_sites"], out["fw_vel"] = sites, vel
    out["fw_dpred"] = jac @ (1 / vel[nearest])
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print("kernels.npz:", len(out), "arrays")


# --------------------------------------------------------------------------- #
# 4. parallel-tempering swap rounds
# --------------------------------------------------------------------------- #
def pt_golden(bb):
    """Tape-driven ``ParallelTempering.on_end_advance_chain`` on 6 polyfit chains."""
    global TAPE
    from bayesbay_b200 import synthetic

    ing = synthetic.polyfit()
    TAPE = Tape(99)
    parameterization, log_likelihood = synthetic.build_problem(ing, bb, reference_forward)
    inv = bb.BayesianInversion(parameterization, log_likelihood, n_chains=6)
    sampler = bb.samplers.ParallelTempering(temperature_max=5, chains_with_unit_temperature=0.4, swap_every=20)
    sampler.initialize(inv.chains)
    sampler.parallel_config = {"n_jobs": 1}
    out = dict(ladder=np.array([c.temperature for c in inv.chains]), rounds=[])
    mis, ldt, t_before, t_after, cur0, cur1 = [], [], [], [], [], []
    for _ in range(5):
        # advance without the swap (call the chains directly), then replay the swap round
        for ch in sampler.chains:
            ch.advance_chain(n_iterations=20, burnin_iterations=0, save_every=100, verbose=False,
                             begin_iteration=sampler.begin_iteration, end_iteration=sampler.end_iteration)
        m_l = [log_likelihood._get_misfit_and_det(ch.current_state) for ch in sampler.chains]
        mis.append([m for m, _ in m_l])
        ldt.append([d for _, d in m_l])
        t_before.append([ch.temperature for ch in sampler.chains])
        cur0.append(TAPE.n)
        sampler.on_end_advance_chain()
        cur1.append(TAPE.n)
        t_after.append([ch.temperature for ch in sampler.chains])
    np.savez_compressed(os.path.join(HERE, "pt_swap.npz"), ladder=out["ladder"], misfit=np.array(mis),
                        logdet=np.array(ldt), t_before=np.array(t_before), t_after=np.array(t_after),
                        cursor0=np.array(cur0), cursor1=np.array(cur1), tape=TAPE.values[: TAPE.n])
    print("pt_swap.npz: ladder", out["ladder"], "swapped rounds",
          [int((np.array(a) != np.array(b)).any()) for a, b in zip(t_before, t_after)])
    # ladders of on_initialize for several chain counts (quirk Q5)
    lad = {}
    for n in (2, 3, 5, 8, 16, 40):
        TAPE = Tape(1)
        chains = [type("C", (), {"temperature": 1.0})() for _ in range(n)]
        bb.samplers.ParallelTempering(temperature_max=7.5, chains_with_unit_temperature=0.4).on_initialize(chains)
        lad[f"n{n}"] = np.array([c.temperature for c in chains])
    np.savez_compressed(os.path.join(HERE, "pt_ladder.npz"), **lad)


def sa_golden(bb):
    """Tape-driven ``SimulatedAnnealing.run`` (samplers/_samplers.py:368-448) on polyfit chains, one chain at a time in
    this process (n_jobs = 1): per iteration the temperature the iteration ran at, the accept count, sigma and the
    coefficients, and the tape position -- the device schedule is replayed against it."""
    global TAPE
    from bayesbay_b200 import synthetic

    ing = synthetic.polyfit()
    n_chains, n_it, burnin = 4, 60, 40
    t_start, frac = 8.0, 0.6
    out = dict(temperature=np.zeros((n_chains, n_it)), accepted=np.zeros((n_chains, n_it), np.int64),
               sigma=np.zeros((n_chains, n_it + 1)), values=np.zeros((n_chains, n_it + 1, 4)),
               cursor=np.zeros((n_chains, n_it + 1), np.int64))
    tapes = []
    for c in range(n_chains):
        TAPE = Tape(500 + c)
        parameterization, log_likelihood = synthetic.build_problem(ing, bb, reference_forward)
        inv = bb.BayesianInversion(parameterization, log_likelihood, n_chains=1)
        chain = inv.chains[0]
        sampler = bb.samplers.SimulatedAnnealing(temperature_start=t_start, cooling_fraction=frac)
        rec = dict(t=[], acc=[], sig=[], val=[], cur=[])

        def snap(ch):
            st = ch.current_state
            rec["sig"].append(st["my_data"].std)
            rec["val"].append([float(np.squeeze(st["my_param_space"][f"m{i}"])) for i in range(4)])
            rec["cur"].append(TAPE.n)

        snap(chain)
        sampler.add_on_begin_iteration(lambda s, ch: None)
        # the temperature an iteration runs at is set by on_begin_iteration; read it right after
        orig_begin = sampler.on_begin_iteration

        def begin(ch, _o=orig_begin):
            _o(ch)
            rec["t"].append(ch.temperature)

        sampler.on_begin_iteration = begin
        sampler.add_on_end_iteration(lambda s, ch: (rec["acc"].append(ch.statistics["n_accepted_models_total"]), snap(ch)))
        inv.run(sampler=sampler, n_iterations=n_it, burnin_iterations=burnin, save_every=10, verbose=False,
                parallel_config={"n_jobs": 1})
        out["temperature"][c] = rec["t"]
        out["accepted"][c] = rec["acc"]
        out["sigma"][c] = rec["sig"]
        out["values"][c] = rec["val"]
        out["cursor"][c] = rec["cur"]
        tapes.append(TAPE.values[: TAPE.n].copy())
        meta = (sampler.cooling_rate, sampler.cooling_iterations)
    L = max(t.size for t in tapes)
    tape = np.zeros((n_chains, L))
    for c, t in enumerate(tapes):
        tape[c, : t.size] = t
    np.savez_compressed(os.path.join(HERE, "sa_schedule.npz"), tape=tape, temperature_start=t_start, cooling_fraction=frac,
                        burnin_iterations=burnin, n_iterations=n_it, cooling_rate=meta[0], cooling_iterations=meta[1], **out)
    print("sa_schedule.npz: temperatures", np.round(out["temperature"][0, ::8], 3), "cooling iterations", meta[1])


# --------------------------------------------------------------------------- #
# 5. long-run statistics with the reference's own RNG (distributional targets)
# --------------------------------------------------------------------------- #
def stats_golden(ing, n_chains, n_iterations, burnin, name, seed):
    """Run in a fresh interpreter (unpatched RNG); see ``_stats_worker``."""
    code = f"""
import sys, random, numpy as np
sys.path[:0] = {[os.path.join(SCRATCH, 'src'), os.path.join(SCRATCH, 'stubs'), REPO, HERE]!r}
import bayesbay as bb
import make_golden as mg
from bayesbay_b200 import synthetic
random.seed({seed}); np.random.seed({seed})
ing = synthetic.{ing}
par, ll = synthetic.build_problem(ing, bb, mg.reference_forward)
inv = bb.BayesianInversion(par, ll, n_chains={n_chains})
inv.run(n_iterations={n_iterations}, burnin_iterations={burnin}, save_every=10, verbose=False,
        parallel_config={{'n_jobs': {min(n_chains, 8)}}})
names = []
prop = []
acc = []
for ch in inv.chains:
    p, a, nm = [], [], []
    for key, v in sorted(ch.statistics['n_proposed_models'].items()):
        if isinstance(v, dict):
            for k2 in sorted(v):
                nm.append(k2); p.append(v[k2]); a.append(ch.statistics['n_accepted_models'][key][k2])
        else:
            nm.append(key); p.append(v); a.append(ch.statistics['n_accepted_models'][key])
    names = nm; prop.append(p); acc.append(a)
res = inv.get_results()
out = dict(names=np.array(names), proposed=np.array(prop), accepted=np.array(acc))
tname = ll.targets[0].name
out['sigma'] = np.array(res[tname + '.std'])
ps_name = list(par.parameter_spaces)[0]
if ps_name + '.n_dimensions' in res:
    out['k'] = np.array(res[ps_name + '.n_dimensions'])
out['dpred_mean'] = np.mean(res[tname + '.dpred'], axis=0)
per_chain = inv.get_results(concatenate_chains=False)
out['dpred_mean_chain'] = np.array([np.mean(v, axis=0) for v in per_chain[tname + '.dpred']])
if 'interp_vel' in per_chain:
    out['field_mean_chain'] = np.array([np.mean(v, axis=0) for v in per_chain['interp_vel']])
for pname in par.parameter_spaces[ps_name].parameters:
    out['chain_mean_' + pname] = np.array([np.mean([np.mean(a) for a in v]) for v in per_chain[ps_name + '.' + pname]])
for pname in par.parameter_spaces[ps_name].parameters:
    vals = res[ps_name + '.' + pname]
    out['mean_' + pname] = np.array([np.mean(v) for v in vals])
np.savez_compressed({os.path.join(HERE, 'stats_' + name + '.npz')!r}, **out)
tot_p = out['proposed'].sum(0); tot_a = out['accepted'].sum(0)
print('stats_{name}:', dict(zip(names, np.round(tot_a / tot_p, 4))), 'sigma mean', out['sigma'].mean())
"""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(SCRATCH, "src"), os.path.join(SCRATCH, "stubs"), REPO, HERE])
    subprocess.run([sys.executable, "-c", code], check=True, env=env)


def main():
    bb = import_reference()
    print("reference version", bb.__version__)
    from bayesbay_b200 import synthetic

    # statistics first (fresh interpreters, unpatched RNG)
    if "--no-stats" not in sys.argv:
        stats_golden("partition_1d()", 8, 60000, 10000, "partition_1d", 11)
        stats_golden("polyfit()", 8, 40000, 5000, "polyfit", 12)
        stats_golden("polyfit(all_gaussian=False)", 8, 40000, 5000, "polyfit_nb", 13)
        stats_golden("tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4, kmax=20)",
                     8, 30000, 5000, "tomo_small", 14)
    install_patches()
    kernels_golden(bb)
    replay(bb, synthetic.tomography_2d(nx=12, ny=10, n_sources_per_side=3, n_receivers=5, kmin=4, kmax=20),
           n_chains=8, n_steps=80, seed0=100, name="tomo_small")
    replay(bb, synthetic.partition_1d(n_data=30, kmin=2, kmax=12), n_chains=8, n_steps=80, seed0=200,
           name="partition_1d")
    replay(bb, synthetic.partition_1d(n_data=30, kmin=2, kmax=12, birth_from="neighbour"), n_chains=4,
           n_steps=80, seed0=250, name="partition_1d_nb")
    replay(bb, synthetic.partition_1d(n_data=30, kmin=2, kmax=12, correlated=True), n_chains=6, n_steps=80, seed0=270,
           name="partition_1d_corr")
    replay(bb, synthetic.partition_1d(n_data=30, kmin=2, kmax=12, birth_from="neighbour", position_dependent=True),
           n_chains=6, n_steps=100, seed0=280, name="partition_1d_posdep")
    replay(bb, synthetic.partition_1d(n_data=30, kmin=2, kmax=12, birth_from="neighbour", prior_kind="laplace"),
           n_chains=6, n_steps=100, seed0=285, name="partition_1d_laplace")
    replay(bb, synthetic.partition_1d_fixed(), n_chains=4, n_steps=60, seed0=290, name="partition_1d_fixed")
    replay(bb, synthetic.transd_polyfit(), n_chains=8, n_steps=100, seed0=295, name="transd_polyfit")
    replay(bb, synthetic.polyfit(), n_chains=8, n_steps=60, seed0=300, name="polyfit")
    replay(bb, synthetic.polyfit(all_gaussian=False), n_chains=4, n_steps=60, seed0=350, name="polyfit_nb")
    pt_golden(bb)
    sa_golden(bb)


if __name__ == "__main__":
    main()
