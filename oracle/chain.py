"""One reversible-jump MCMC step, restated from the reference chain loop.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Citations relative to ``/root/reference``.

The problem is described by a *spec dict* (the same plain-data format
``bayesbay_b200._compile.compile_problem`` emits)::

    {"spaces": [{"name", "kind": "plain"|"voronoi1d"|"voronoi2d", "ndim", "trans_d",
                 "kmin", "kmax", "kinit_max", "vmin": [..], "vmax": [..], "site_std": [..],
                 "birth_from": "prior"|"neighbour",
                 "params": [{"name", "kind": "uniform"|"gaussian"|"laplace", "a", "b",
                             "perturb_std", "perturb_std_birth"}],
                 "weights": {"birth", "death", "values", "sites"}}],
     "targets": [{"name", "dobs", "hierarchical", "std_min", "std_max", "std_perturb_std",
                  "cov_inv": None | float | 1-D array,
                  "forward": {"kind": "ray2d", "space", "param", "transform", "indptr", "indices",
                              "data", "points"}
                           | {"kind": "lookup1d", "space", "param", "x"}
                           | {"kind": "linear", "space", "params": [..], "G"}}]}
"""
from __future__ import annotations

import math
from bisect import bisect_left
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import kernels as K
from . import rng as R

INNER = ("birth", "death", "values", "sites")
MOVE_BIRTH, MOVE_DEATH, MOVE_VALUES, MOVE_SITES = range(4)


# --------------------------------------------------------------------------- #
# state
# --------------------------------------------------------------------------- #
@dataclass
class SpaceState:
    k: int
    sites: Optional[np.ndarray]            # (k, 2) | (k,) | None
    values: List[np.ndarray]               # one (k,) array per parameter

    def copy(self):
        return SpaceState(self.k, None if self.sites is None else self.sites.copy(),
                          [v.copy() for v in self.values])


@dataclass
class ChainState:
    spaces: List[SpaceState]
    sigma: List[Optional[float]]
    T: float = 1.0
    dpred: List[Optional[np.ndarray]] = field(default_factory=list)
    corr: List[Optional[float]] = field(default_factory=list)  # noise correlation per target (None: uncorrelated)

    def __post_init__(self):
        if not self.corr:
            self.corr = [None] * len(self.sigma)

    def copy(self):
        return ChainState([s.copy() for s in self.spaces], list(self.sigma), self.T,
                          [None if d is None else d for d in self.dpred], list(self.corr))


def outer_moves(spec):
    """The outer perturbation list and weights of ``BayesianInversion._init_perturbation_funcs``
    (src/bayesbay/_bayes_inversion.py:458-466): one ``ParamSpacePerturbation`` per space, weight =
    sum of its inner weights (``_parameter_space.py:324-352``; Voronoi adds the site move,
    ``discretization/_voronoi.py:402-430``), then one ``NoisePerturbation`` of weight 1 if any
    target is hierarchical (``likelihood/_log_likelihood.py:293-306``)."""
    moves = []
    weights = []
    for i, sp in enumerate(spec["spaces"]):
        inner = [(j, sp["weights"][name]) for j, name in enumerate(INNER) if sp["weights"][name] > 0]
        moves.append(("space", i, inner))
        weights.append(sp.get("w_outer") or sum(w for _, w in inner))
    # (custom weights: BaseBayesianInversion.set_perturbation_funcs, _bayes_inversion.py:117-171)
    w_noise = spec.get("w_noise")
    if any(t["hierarchical"] for t in spec["targets"]) and w_noise != 0:
        moves.append(("noise", None, None))
        weights.append(1 if w_noise is None else w_noise)
    return moves, weights


def n_move_types(spec):
    return 4 * len(spec["spaces"]) + 1


# --------------------------------------------------------------------------- #
# I1: initialisation
# --------------------------------------------------------------------------- #
def _prior_sample(p, s, position=None):
    """``UniformPrior.sample`` (prior/_prior.py:390-394), ``GaussianPrior.sample`` (:571-575),
    ``LaplacePrior.sample``, hyper-parameters taken at ``position``."""
    p = K.prior_at(p, position)
    if p["kind"] == "uniform":
        return R.draw_uniform(s, p["a"], p["b"])
    if p["kind"] == "gaussian":
        return R.draw_normal(s, p["a"], p["b"])
    return R.draw_laplace(s, p["a"], p["b"])


def _sample_site(sp, s):
    """``Voronoi.sample_site`` (discretization/_voronoi.py:118-120), 1-D :536-538."""
    if sp["ndim"] == 1:
        return R.draw_uniform(s, sp["vmin"][0], sp["vmax"][0])
    return np.array([R.draw_uniform(s, sp["vmin"][d], sp["vmax"][d]) for d in range(sp["ndim"])])


def init_state(spec, s, temperature=1.0):
    """``Voronoi._initialize`` (discretization/_voronoi.py:151-169; 1-D :551-564),
    ``ParameterSpace._initialize`` (parameterization/_parameter_space.py:146-160),
    ``Prior.initialize`` (prior/_prior.py:396-398, :577-597), ``Target.initialize``
    (likelihood/_target.py:115-134).  Draw order: k, then all sites, then each parameter's k values,
    then each hierarchical target's sigma."""
    spaces = []
    for sp in spec["spaces"]:
        k = R.draw_randint(s, sp["kmin"], sp["kinit_max"]) if sp["trans_d"] else sp["kmin"]
        sites = None
        if sp["ndim"] == 1:
            sites = np.sort(np.array([_sample_site(sp, s) for _ in range(k)]))
        elif sp["ndim"] == 2:
            sites = np.array([_sample_site(sp, s) for _ in range(k)]).reshape(k, 2)
        values = [np.array([_prior_sample(p, s, None if sites is None or sp["ndim"] != 1 else sites[j]) for j in range(k)])
                  for p in sp["params"]]
        spaces.append(SpaceState(k, sites, values))
    sigma, corr = [], []
    for t in spec["targets"]:
        sigma.append(R.draw_uniform(s, t["std_min"], t["std_max"]) if t["hierarchical"] else None)
        corr.append(R.draw_uniform(s, t["corr_min"], t["corr_max"]) if t.get("correlated") else None)
    st = ChainState(spaces, sigma, temperature, [None] * len(spec["targets"]), corr)
    evaluate(spec, st)
    return st


# --------------------------------------------------------------------------- #
# forward + likelihood
# --------------------------------------------------------------------------- #
def forward_target(t, st):
    fwd = t["forward"]
    sp = st.spaces[fwd["space"]]
    if fwd["kind"] == "ray2d":
        return K.tomography_forward(fwd, sp.sites, sp.values[fwd["param"]])[0]
    if fwd["kind"] == "lookup1d":
        return K.lookup1d_forward(fwd["x"], sp.sites, sp.values[fwd["param"]])
    if fwd["kind"] == "linear":
        m = np.concatenate([sp.values[p] for p in fwd["params"]])
        return K.linear_forward(fwd["G"], m)
    if fwd["kind"] == "polynomial":  # docs/source/tutorials/03_transd_polyfit.ipynb cell 5
        m = sp.values[fwd["param"]]
        return np.vander(fwd["x"], len(m), True) @ m
    raise ValueError(fwd["kind"])


def evaluate(spec, st):
    """Fill the per-target d_pred cache (``"{target}.dpred"``, likelihood/_log_likelihood.py:311-320)."""
    for i, t in enumerate(spec["targets"]):
        if st.dpred[i] is None:
            st.dpred[i] = forward_target(t, st)


def misfit_and_logdet(spec, st):
    """``LogLikelihood._get_misfit_and_det`` summed over targets (likelihood/_log_likelihood.py:308-332)."""
    misfit = 0.0
    logdet = 0.0
    for i, t in enumerate(spec["targets"]):
        m, ld = K.misfit_and_logdet(st.dpred[i], t["dobs"], st.sigma[i], t["cov_inv"], st.corr[i])
        misfit += m
        if t["hierarchical"]:
            logdet += ld
    return misfit, logdet


# --------------------------------------------------------------------------- #
# P1-P7: proposals
# --------------------------------------------------------------------------- #
def _nearest(sp, sites, pos):
    if sp["ndim"] == 1:
        return K.nearest_neighbour_1d(float(pos), sites)
    return K.nearest_neighbour_2d(sites, pos)


def _perturb_site(sp, site, s):
    """``Voronoi._perturb_site`` (discretization/_voronoi.py:171-188; 1-D :566-571): resample until
    inside the box (quirk Q7: truncated proposal treated as symmetric)."""
    if sp["ndim"] == 1:
        while True:
            new = site + R.draw_normal(s, 0.0, sp["site_std"][0])
            if sp["vmin"][0] <= new <= sp["vmax"][0]:
                return new
    while True:
        dev = np.array([R.draw_normal(s, 0.0, sp["site_std"][d]) for d in range(sp["ndim"])])
        new = site + dev
        if all((new >= np.asarray(sp["vmin"])) & (new <= np.asarray(sp["vmax"]))):
            return new


def _move_values(sp, ss, s):
    """``ParamPerturbation.perturb_param_space_state`` (perturbations/_param_values.py:60-84): one
    index, every parameter perturbed there."""
    idx = R.draw_randint(s, 0, ss.k - 1)
    new = ss.copy()
    lpr = 0.0
    for j, p in enumerate(sp["params"]):
        p = K.prior_at(p, ss.sites[idx] if sp["ndim"] == 1 else None)
        old = float(ss.values[j][idx])
        val = old + R.draw_normal(s, 0.0, p["perturb_std"])
        lpr += K.prior_value_ratio(p, old, val)
        new.values[j][idx] = val
    return new, lpr


def _move_site(sp, ss, s):
    """``Voronoi.perturb_value`` (discretization/_voronoi.py:190-232); 1-D re-sorts sites and
    permutes the values (:573-591).  Prior ratio is 0 for position-independent priors."""
    idx = R.draw_randint(s, 0, ss.k - 1)
    new = ss.copy()
    new.sites[idx] = _perturb_site(sp, ss.sites[idx], s)
    lpr = 0.0
    for j, p in enumerate(sp["params"]):
        if p.get("table") is not None:  # position-dependent prior: log p(v; new site) - log p(v; old site), :220-226
            v = float(ss.values[j][idx])
            lpr += K.prior_log_pdf(K.prior_at(p, new.sites[idx]), v) - K.prior_log_pdf(K.prior_at(p, ss.sites[idx]), v)
    if sp["ndim"] == 1:
        order = np.argsort(new.sites, kind="stable")
        new.sites = new.sites[order]
        new.values = [v[order] for v in new.values]
    return new, lpr


def _birth(sp, ss, s):
    """Voronoi: ``Voronoi.birth`` (discretization/_voronoi.py:234-339; 1-D :593-611) with
    ``Discretization._initialize_newborn_params`` (discretization/_discretization.py:268-375).
    Plain space: ``ParameterSpace.birth`` (parameterization/_parameter_space.py:197-275)."""
    if ss.k == sp["kmax"]:
        return ss, -math.inf
    if sp["kind"] == "plain":
        i_ins = R.draw_randint(s, 0, ss.k)
        vals = [K.insert_1d(v, i_ins, _prior_sample(p, s)) for p, v in zip(sp["params"], ss.values)]
        return SpaceState(ss.k + 1, None, vals), 0.0
    site = _sample_site(sp, s)
    lpr = 0.0
    born = []
    pos = site if sp["ndim"] == 1 else None
    if sp["birth_from"] == "prior":
        born = [_prior_sample(p, s, pos) for p in sp["params"]]
    else:
        i_near = _nearest(sp, ss.sites, site)
        for j, p in enumerate(sp["params"]):
            p = K.prior_at(p, pos)
            theta = p["perturb_std_birth"]
            old = float(ss.values[j][i_near])
            val = old + R.draw_normal(s, 0.0, theta)
            born.append(val)
            lpr += K.prior_log_pdf(p, val) + math.log(theta * K.SQRT_TWO_PI) + (val - old) ** 2 / (2 * theta ** 2)
    if sp["ndim"] == 1:
        i_ins = bisect_left(ss.sites, site)
        sites = K.insert_1d(ss.sites, i_ins, site)
        vals = [K.insert_1d(v, i_ins, b) for v, b in zip(ss.values, born)]
    else:
        sites = np.vstack((ss.sites, site))
        vals = [np.append(v, b) for v, b in zip(ss.values, born)]
    return SpaceState(ss.k + 1, sites, vals), lpr


def _death(sp, ss, s):
    """``Voronoi.death`` (discretization/_voronoi.py:341-385; 1-D :613-625) with
    ``_log_prob_death_parameters`` (discretization/_discretization.py:435-465); plain space:
    ``ParameterSpace.death`` (parameterization/_parameter_space.py:277-322)."""
    if ss.k == sp["kmin"]:
        return ss, -math.inf
    i_rm = R.draw_randint(s, 0, ss.k - 1)
    vals = [np.delete(v, i_rm) for v in ss.values]
    if sp["kind"] == "plain":
        return SpaceState(ss.k - 1, None, vals), 0.0
    sites = np.delete(ss.sites, i_rm, axis=0)
    lpr = 0.0
    if sp["birth_from"] == "neighbour":
        pos = ss.sites[i_rm]
        i_near = _nearest(sp, sites, pos)
        for j, p in enumerate(sp["params"]):
            p = K.prior_at(p, pos if sp["ndim"] == 1 else None)
            theta = p["perturb_std_birth"]
            gone = float(ss.values[j][i_rm])
            near = float(vals[j][i_near])
            lpr -= K.prior_log_pdf(p, gone)
            lpr -= math.log(theta * K.SQRT_TWO_PI) + (gone - near) ** 2 / (2 * theta ** 2)
    return SpaceState(ss.k - 1, sites, vals), lpr


def _move_noise(spec, st, s):
    """``NoisePerturbation.perturb`` (perturbations/_data_noise.py:21-77): every hierarchical
    target's sigma (then its correlation, if the noise is correlated) random-walks, resampled until inside
    its range; ratio 0."""
    sigma, corr = list(st.sigma), list(st.corr)
    for i, t in enumerate(spec["targets"]):
        if not t["hierarchical"]:
            continue
        while True:
            new = st.sigma[i] + R.draw_normal(s, 0.0, t["std_perturb_std"])
            if new < t["std_min"] or new > t["std_max"]:
                continue
            break
        sigma[i] = new
        if t.get("correlated"):
            while True:
                new = st.corr[i] + R.draw_normal(s, 0.0, t["corr_perturb_std"])
                if new < t["corr_min"] or new > t["corr_max"]:
                    continue
                break
            corr[i] = new
    return sigma, corr


def propose(spec, st, s):
    """Outer + inner weighted choice (``_markov_chain.py:208-210``,
    ``perturbations/_param_space.py:79-92``) and the chosen proposal.
    Returns ``(new_state, log_prob_ratio, move_id)`` with ``move_id = 4*space + inner`` or
    ``4*n_spaces`` for the noise move."""
    moves, weights = outer_moves(spec)
    kind, isp, inner = moves[R.draw_choice(s, weights)]
    new = st.copy()
    if kind == "noise":
        new.sigma, new.corr = _move_noise(spec, st, s)
        # quirk Q1: the reference re-runs the forward here; the result is identical, so keep d_pred
        return new, 0.0, 4 * len(spec["spaces"])
    sp = spec["spaces"][isp]
    j = inner[R.draw_choice(s, [w for _, w in inner])][0]
    ss = st.spaces[isp]
    if j == MOVE_BIRTH:
        ns, lpr = _birth(sp, ss, s)
    elif j == MOVE_DEATH:
        ns, lpr = _death(sp, ss, s)
    elif j == MOVE_VALUES:
        ns, lpr = _move_values(sp, ss, s)
    else:
        ns, lpr = _move_site(sp, ss, s)
    new.spaces[isp] = ns
    for i, t in enumerate(spec["targets"]):
        if t["forward"]["space"] == isp:
            new.dpred[i] = None
    return new, lpr, 4 * isp + j


def step(spec, st, s):
    """``BaseMarkovChain._next_iteration`` (src/bayesbay/_markov_chain.py:204-247).
    Returns ``(state_after, info)``."""
    new, lpr, move = propose(spec, st, s)
    if math.isinf(lpr):
        llr = 0.0
    else:
        evaluate(spec, new)
        om, old = misfit_and_logdet(spec, st)
        nm, nld = misfit_and_logdet(spec, new)
        llr = K.log_like_ratio(om, old, nm, nld, st.T)
    log_u = math.log(R.draw_unit(s))
    accepted = (lpr + llr) > log_u
    info = dict(move=move, log_prob_ratio=lpr, log_like_ratio=llr, accepted=bool(accepted), n_draws=s.n)
    return (new if accepted else st), info


# --------------------------------------------------------------------------- #
# T1: parallel tempering
# --------------------------------------------------------------------------- #
def pt_ladder(n_chains, temperature_max=5.0, chains_with_unit_temperature=0.4):
    """``ParallelTempering.on_initialize`` (src/bayesbay/samplers/_samplers.py:315-326), quirk Q5."""
    ones = np.ones(max(2, int(n_chains * chains_with_unit_temperature)) - 1)
    size = n_chains - ones.size
    return np.concatenate((ones, np.geomspace(1, temperature_max, size)))


def pt_swap_round(misfit, logdet, temps, s):
    """``ParallelTempering.on_end_advance_chain`` (src/bayesbay/samplers/_samplers.py:334-342):
    ``n_chains`` sequential random-pair attempts; quirk Q2 (the ratio is already divided by T1)
    is reproduced.  ``misfit``/``logdet`` are per-chain totals with each chain's own sigma.
    Mutates and returns ``temps``; also returns the number of accepted swaps."""
    n = len(temps)
    n_acc = 0
    for _ in range(n):
        a, b = R.draw_pair(s, n)
        t1, t2 = temps[a], temps[b]
        llr = (logdet[a] - logdet[b] + misfit[a] - misfit[b]) / 2 / t1
        prob = (1 / t1 - 1 / t2) * llr
        if prob > math.log(R.draw_unit(s)):
            temps[a], temps[b] = t2, t1
            n_acc += 1
    return temps, n_acc


# --------------------------------------------------------------------------- #
# chain driver (used by tests and by the CPU baseline)
# --------------------------------------------------------------------------- #
def run_chain(spec, st, n_steps, make_stream, iter0=0, on_step=None):
    """Advance one chain ``n_steps`` iterations; ``make_stream(iteration)`` gives the step's
    uniform stream.  Returns the final state and per-move-type ``(proposed, accepted)`` counts."""
    nm = n_move_types(spec)
    proposed = np.zeros(nm, dtype=np.int64)
    accepted = np.zeros(nm, dtype=np.int64)
    for i in range(n_steps):
        st, info = step(spec, st, make_stream(iter0 + i))
        proposed[info["move"]] += 1
        accepted[info["move"]] += info["accepted"]
        if on_step is not None:
            on_step(iter0 + i, st, info)
    return st, proposed, accepted
