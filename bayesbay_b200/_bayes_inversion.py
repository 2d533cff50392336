"""``BayesianInversion`` facade (reference src/bayesbay/_bayes_inversion.py) on the CUDA engine.

Same constructor and ``run`` / ``get_results`` signatures.  ``run`` does exactly what the reference
does with a sampler (``_bayes_inversion.py:236-249``): ``sampler.initialize(chains)``; set
``sampler.parallel_config``; ``self._chains = sampler.run(...)``.  The chains live on the device from
construction on (starting states are drawn by the initialisation kernel, or taken from
``walkers_starting_states``).

Multi-GPU: under ``torchrun`` (``torch.distributed`` initialised, one rank per GPU) ``n_chains`` is the
job-wide number of chains and each rank owns a contiguous shard; the only cross-rank traffic is the
parallel-tempering all-gather of three doubles per chain.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Dict, List, Union

from . import _compile, _dist
from ._batch import ChainBatch
from ._markov_chain import BaseMarkovChain, MarkovChain
from ._state import State
from .likelihood import LogLikelihood
from .parameterization import Parameterization
from .samplers import Sampler, VanillaSampler

__all__ = ["BaseBayesianInversion", "BayesianInversion"]


class BaseBayesianInversion:
    """Low-level interface of the reference (``_bayes_inversion.py:22-398``), which takes arbitrary
    Python perturbation / likelihood callables.  Those cannot run inside CUDA kernels and there is no
    CPU fallback, so constructing it directly raises; :class:`BayesianInversion` is the supported entry
    point and inherits ``run`` / ``get_results`` from here."""

    def __init__(self, walkers_starting_states=None, perturbation_funcs=None, perturbation_weights=None,
                 log_like_ratio_func=None, log_like_func=None, n_chains: int = 10, save_dpred: bool = True):
        raise TypeError(
            "BaseBayesianInversion drives user-defined Python perturbation and likelihood callables, which cannot "
            "run on the device (no CPU fallback). Describe the problem with Parameterization + LogLikelihood and use "
            "BayesianInversion.")

    @property
    def chains(self) -> List[BaseMarkovChain]:
        return self._chains

    def run(self, sampler: Sampler = None, n_iterations: int = 1000, burnin_iterations: int = 0, save_every: int = 100,
            verbose: bool = True, print_every: int = 100, parallel_config: dict = None):
        """Run the inversion (reference ``_bayes_inversion.py:198-249``).  ``parallel_config`` is kept for
        signature compatibility; joblib keys (``n_jobs``, ``backend``) are ignored."""
        if sampler is None:
            sampler = VanillaSampler()
        sampler.initialize(self.chains)
        _parallel_config = {"n_jobs": len(self.chains), "backend": "cuda"}
        if parallel_config is not None:
            _parallel_config.update(parallel_config)
        sampler.parallel_config = _parallel_config
        self._chains = sampler.run(n_iterations=n_iterations, burnin_iterations=burnin_iterations,
                                   save_every=save_every, verbose=verbose, print_every=print_every)
        # the samples are in host memory when run() returns (save points travel while the chains keep stepping); the
        # per-chain views (saved_states, statistics, current_state) are filled when they are first read
        self.batch.flush_saves()
        self._mark_dirty()

    def get_results(self, keys: Union[str, List[str]] = None, concatenate_chains: bool = True):
        """Saved samples of this rank's chains (reference ``_bayes_inversion.py:251-326``)."""
        if concatenate_chains and getattr(self, "batch", None) is not None:
            fast = self.batch.results([keys] if isinstance(keys, str) else keys)
            if fast is not None:
                return fast
        return self.get_results_from_chains(self.chains, keys, concatenate_chains)

    @staticmethod
    def get_results_from_chains(chains, keys=None, concatenate_chains=True) -> Dict[str, list]:
        if not isinstance(chains, list):
            chains = [chains]
        if not all(isinstance(c, BaseMarkovChain) for c in chains):
            raise TypeError("`chains` should be a list of Markov chains (i.e. instances of BaseMarkovChain or MarkovChain)")
        if chains and getattr(chains[0], "_batch_owner", None) is not None:
            chains[0]._batch_owner._sync_chains()
        if isinstance(keys, str):
            keys = [keys]
        results = defaultdict(list)
        for chain in chains:
            for key, saved_values in chain.saved_states.items():
                if keys is None or key in keys:
                    if concatenate_chains and isinstance(saved_values, list):
                        results[key].extend(saved_values)
                    else:
                        results[key].append(saved_values)
        return dict(results)


class BayesianInversion(BaseBayesianInversion):
    """``BayesianInversion(parameterization, log_likelihood, n_chains=10, walkers_starting_states=None,
    save_dpred=True)`` -- reference ``_bayes_inversion.py:329-488``.

    Extra keyword-only arguments (all optional): ``device`` (default ``cuda:$LOCAL_RANK``) and ``seed``
    (Philox key; default from :func:`bayesbay_b200.seed` or OS entropy)."""

    def __init__(self, parameterization: Parameterization, log_likelihood: LogLikelihood, n_chains: int = 10,
                 walkers_starting_states: List[State] = None, save_dpred: bool = True, *, device=None, seed=None):
        if not isinstance(log_likelihood, LogLikelihood):
            raise TypeError("`log_likelihood` should be an instance of `bayesbay.likelihood.LogLikelihood`")
        if walkers_starting_states is not None and n_chains != len(walkers_starting_states):
            raise ValueError(
                f"Length of `walkers_starting_states` ({len(walkers_starting_states)})"
                f" does not match `n_chains` ({n_chains}).")
        self.parameterization = parameterization
        self.log_likelihood = log_likelihood
        self.n_chains = n_chains
        self.save_dpred = save_dpred
        self.walkers_starting_states = walkers_starting_states
        self.perturbation_funcs = parameterization.perturbation_funcs + log_likelihood.perturbation_funcs
        self.perturbation_weights = parameterization.perturbation_weights + log_likelihood.perturbation_weights
        self.spec = _compile.compile_problem(parameterization, log_likelihood)
        # chain sharding over ranks (one process per GPU)
        self.rank, self.world_size = _dist.world()
        self.n_chains_global = n_chains
        self.chain_id0, n_local = _dist.shard(n_chains)
        if device is None:
            device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
        if seed is None:
            from ._batch import next_seed

            seed = next_seed()
        seed = _dist.broadcast_seed(seed, device)
        starts = None
        if walkers_starting_states is not None:
            starts = walkers_starting_states[self.chain_id0: self.chain_id0 + n_local]
        self.batch = ChainBatch(self.spec, n_local, device=device, seed=seed, chain_id0=self.chain_id0,
                                starting_states=starts, save_dpred=save_dpred)
        self.batch.seed_global = seed
        self._chains = []
        for c in range(n_local):
            ch = MarkovChain(id=self.chain_id0 + c, batch=self.batch, index=c, parameterization=parameterization,
                             log_likelihood=log_likelihood, perturbation_funcs=self.perturbation_funcs,
                             perturbation_weights=self.perturbation_weights, save_dpred=save_dpred)
            ch._batch_owner = self
            self._chains.append(ch)
        self._dirty = True
        self._n_saved_synced = 0
        self.log_likelihood.add_targets_observer(self)

    # ---- device <-> chain views -------------------------------------------
    def _mark_dirty(self):
        self._dirty = True
        for ch in self._chains:
            ch._current_state = None

    def _sync_chains(self, states: bool = True):
        """Refresh the per-chain views (statistics, saved states, temperature; current state on request)."""
        if self._dirty:
            self._dirty = False  # first: the chain properties below consult it
            stats = self.batch.statistics()
            temps = self.batch.temperatures()
            start = self._n_saved_synced
            for c, ch in enumerate(self._chains):
                ch._statistics = stats[c]
                ch._temperature = float(temps[c])
                for key, vals in self.batch.saved_states_of(c, start).items():
                    ch._saved_states[key].extend(vals)
            self._n_saved_synced = len(self.batch.saved)
        if states and self._chains and self._chains[0]._current_state is None:
            for ch, st in zip(self._chains, self.batch.current_states()):
                ch._current_state = st

    def get_field_statistics(self, name: str) -> dict:
        """Posterior mean / standard deviation map of a field a forward operator saves (``save_field_as``),
        accumulated on the device over this rank's saved samples (no per-sample storage needed)."""
        return self.batch.field_statistics(name)

    # ---- checkpoint / resume (the reference has none) ------------------------
    def save_checkpoint(self, path):
        """Write this rank's chains (device state, iteration counter, samples saved so far) to ``path``."""
        import torch

        torch.save(self.batch.state_dict(), path)

    def load_checkpoint(self, path):
        """Continue bit-identically from :meth:`save_checkpoint` (same problem, n_chains, seed and sharding).  The
        NEXT ``run`` continues the tempering state of the checkpoint: ``ParallelTempering`` keeps the per-chain
        temperatures as the swap rounds left them (instead of laying the initial ladder again) and
        ``SimulatedAnnealing`` keeps its schedule position.  The file is a ``torch.save`` pickle of tensors, numpy
        arrays and ints: load checkpoints from trusted sources only."""
        import torch

        self.batch.load_state_dict(torch.load(path, weights_only=False))
        self.batch.resumed = True
        self._n_saved_synced = 0
        for ch in self._chains:
            ch._saved_states.clear()
        self._mark_dirty()

    def all_gather(self, local):
        """[C_local][3] -> [C_total][3] in global chain order (NCCL all-gather over NVLink when sharded)."""
        return _dist.all_gather(local)

    def update_log_likelihood_targets(self, targets):
        raise TypeError("targets cannot be added after the chains were placed on the device; build a new inversion")

    def set_perturbation_funcs(self, perturbation_funcs, perturbation_weights=None):
        """Custom perturbation list and weights (reference ``_bayes_inversion.py:117-171``).  The device engine runs
        the library's own perturbations only, so the list must be made of THIS inversion's perturbation objects: the
        ``ParamSpacePerturbation`` of a space (a new weight for the space as a whole), or its inner birth / death /
        value / site perturbations one by one (``inversion.perturbation_funcs[0].perturbation_funcs``, as the
        reference's tests 11 and 19 do -- each then gets its own weight and its own statistics entry), and the
        ``NoisePerturbation`` (left out: the noise level is never perturbed).  Like the reference, a death weight
        that differs from its birth weight is reset to the birth weight.  Call it before the first ``run``."""
        import warnings

        from .perturbations import (BirthPerturbation, DeathPerturbation, NoisePerturbation, ParamPerturbation,
                                    ParamSpacePerturbation)

        funcs = list(perturbation_funcs)
        weights = [1] * len(funcs) if perturbation_weights is None else list(perturbation_weights)
        assert len(funcs) == len(weights), (
            "`perturbation_funcs` should have the same length of "
            f"`perturbation_weights`: {len(funcs)} != {len(weights)}")
        spec = dict(self.spec, spaces=[dict(sp) for sp in self.spec["spaces"]])
        index = {sp["name"]: i for i, sp in enumerate(spec["spaces"])}
        kinds = ("birth", "death", "values", "sites")
        whole, inner = {}, {}
        w_noise = 0.0
        for f, w in zip(funcs, weights):
            if not (w > 0):
                raise ValueError("perturbation weights must be positive")
            if isinstance(f, ParamSpacePerturbation):
                whole[f.param_space_name] = whole.get(f.param_space_name, 0.0) + float(w)
            elif isinstance(f, (BirthPerturbation, DeathPerturbation, ParamPerturbation)):
                d = inner.setdefault(f.param_space_name, {})
                if kinds[f.move_kind] in d:
                    raise ValueError(f"{f.__name__} is listed twice")
                d[kinds[f.move_kind]] = float(w)
            elif isinstance(f, NoisePerturbation):
                w_noise += float(w)
            else:
                raise TypeError(f"{f!r}: only the perturbation objects of this inversion run on the device (Python "
                                "callables cannot run in a CUDA kernel and there is no CPU fallback)")
        if self.batch.iteration > 0:
            raise RuntimeError("set_perturbation_funcs must be called before the chains have advanced")
        for name in set(whole) | set(inner):
            if name not in index:
                raise ValueError(f"unknown parameter space {name!r}")
            if name in whole and name in inner:
                raise ValueError(f"parameter space {name!r}: list either its ParamSpacePerturbation or its inner perturbations")
        for sp in spec["spaces"]:
            name = sp["name"]
            if name in whole:
                sp["w_outer"], sp["flat"] = whole[name], False
            elif name in inner:
                d = inner[name]
                if ("birth" in d) != ("death" in d):
                    raise ValueError("there should be exactly one birth and one death perturbation function for each "
                                     f"trans-dimensional parameter space, but {name} has only one of them")
                if "birth" in d and d["birth"] != d["death"]:
                    warnings.warn(f"weights for the birth and death perturbations of {name} are different: {d['birth']} != "
                                  f"{d['death']}. We will default the death perturbation weight to be the same as birth weight")
                    d["death"] = d["birth"]
                sp["weights"] = {k: d.get(k, 0.0) for k in kinds}
                sp["w_outer"], sp["flat"] = None, True
            else:
                raise ValueError(f"parameter space {name!r} has no perturbation in the list (its state would never move)")
        spec["w_noise"] = w_noise if any(t["hierarchical"] for t in spec["targets"]) else None
        self.spec = spec
        self.perturbation_funcs, self.perturbation_weights = funcs, weights
        old = self.batch
        self.batch = ChainBatch(spec, old.n_chains, device=old.engine.device, seed=old.seed, chain_id0=old.chain_id0,
                                starting_states=old.current_states(), save_dpred=old.save_dpred)
        self.batch.seed_global = getattr(old, "seed_global", old.seed)
        for ch in self._chains:
            ch._batch = self.batch
            ch.set_perturbation_funcs(funcs, weights)
        self._mark_dirty()

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(parameterization={self.parameterization}, "
                f"log_likelihood={self.log_likelihood}, n_chains={self.n_chains}, save_dpred={self.save_dpred})")
