"""Plug-in for the UNMODIFIED reference: ``Sampler`` subclasses whose ``run()`` advances the reference's chains on
the CUDA engine instead of through ``Sampler.advance_chain``'s joblib / loky process pool.

The reference's seam: ``BaseBayesianInversion.run`` does three things with a sampler
(src/bayesbay/_bayes_inversion.py:236-249) -- ``sampler.initialize(self.chains)``, sets ``sampler.parallel_config``,
``self._chains = sampler.run(n_iterations=, burnin_iterations=, save_every=, verbose=, print_every=)`` -- and
``Sampler.run`` is abstract (samplers/_samplers.py:152-162).  So, with the reference imported as ``bayesbay``::

    import bayesbay
    from bayesbay_b200.forward import RayTomographyForward
    from bayesbay_b200.plugin import cuda_samplers

    CudaVanillaSampler, CudaParallelTempering = cuda_samplers(bayesbay)
    log_likelihood = bayesbay.likelihood.LogLikelihood(targets=target, fwd_functions=RayTomographyForward(...))
    inversion = bayesbay.BayesianInversion(parameterization, log_likelihood, n_chains=1024)
    inversion.run(sampler=CudaVanillaSampler(), n_iterations=..., burnin_iterations=..., save_every=...)
    results = inversion.get_results()          # the reference's own get_results_from_chains

The forward functions must be the device operator descriptors of :mod:`bayesbay_b200.forward` (they are valid
reference forwards: the reference calls them on the host while it initialises its chains); anything else raises
``TypeError`` when the run starts -- there is no CPU fallback.  The chains start from the states the reference
initialised, come back with ``saved_states`` / ``statistics`` / ``current_state`` / ``temperature`` filled the way
``MarkovChain.advance_chain`` fills them (src/bayesbay/_markov_chain.py:100-157, 249-297), and the reference's
``get_results`` / ``get_results_from_chains`` work on them unchanged.  Per-iteration Python callbacks
(``add_on_begin_iteration`` / ``add_on_end_iteration``) cannot run inside the batched step and are rejected.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

from . import _compile
from ._batch import ChainBatch

__all__ = ["cuda_samplers"]


def _merge_statistics(chain, stats: dict):
    """Add one device run's counters to a reference chain's ``statistics`` (``_init_statistics`` /
    ``_save_statistics``, src/bayesbay/_markov_chain.py:106-157)."""
    st = chain._statistics
    for key in ("n_proposed_models", "n_accepted_models"):
        for outer, val in stats[key].items():
            if isinstance(val, dict):
                for inner, n in val.items():
                    st[key][outer][inner] += n
            else:
                prev = st[key].get(outer, 0)
                st[key][outer] = (prev if not isinstance(prev, dict) else 0) + val
    st["n_proposed_models_total"] += stats["n_proposed_models_total"]
    st["n_accepted_models_total"] += stats["n_accepted_models_total"]
    for k, v in stats["exceptions"].items():
        st["exceptions"][k] += v


def cuda_samplers(bayesbay, device=None, seed=None):
    """``(CudaVanillaSampler, CudaParallelTempering)`` as subclasses of ``bayesbay.samplers.Sampler`` -- the
    reference module is passed in, nothing of it is imported or copied here."""
    Sampler = bayesbay.samplers.Sampler

    class _CudaSampler(Sampler):
        def __init__(self, device=device, seed=seed):
            super().__init__()
            self._device = device
            self._seed = seed
            self.batch = None

        # the reference's per-chain hooks: nothing to do for the built-in behaviours reproduced on the device
        def on_initialize(self, chains):
            pass

        def on_begin_iteration(self, chain):
            pass

        def on_end_iteration(self, chain):
            pass

        def on_end_advance_chain(self):
            pass

        def add_on_begin_iteration(self, func):
            raise TypeError("per-iteration Python callbacks cannot run inside the batched CUDA step")

        add_on_end_iteration = add_on_begin_iteration

        # ---- chains <-> device batch --------------------------------------------------------------------
        def _open(self):
            chains = self.chains
            first = chains[0]
            if not hasattr(first, "parameterization") or not hasattr(first, "log_likelihood"):
                raise TypeError("the CUDA samplers need MarkovChain objects built by bayesbay.BayesianInversion "
                                "(BaseBayesianInversion's perturbation / likelihood callables cannot run on the device)")
            spec = _compile.compile_problem(first.parameterization, first.log_likelihood)
            self.batch = ChainBatch(spec, len(chains), device=self._device, seed=self._seed,
                                    starting_states=[ch.current_state for ch in chains],
                                    save_dpred=bool(getattr(first, "save_dpred", True)))
            self.batch.iteration = int(first.ith_iteration)  # the save schedule counts a chain's whole life
            self.batch.seed_global = self.batch.seed
            self.batch.set_temperatures([float(ch.temperature) for ch in chains])
            self._saved_from = 0

        def _close(self):
            """Hand the run's samples, counters and final states to the reference's chain objects."""
            batch = self.batch
            states = batch.current_states(state_cls=bayesbay.State, ps_cls=bayesbay.ParameterSpaceState,
                                          noise_cls=bayesbay.DataNoiseState)
            stats = batch.statistics()
            temps = batch.temperatures()
            for c, ch in enumerate(self.chains):
                if ch.saved_states is None:
                    ch._saved_states = defaultdict(list)
                for key, values in batch.saved_states_of(c, self._saved_from).items():
                    ch.saved_states[key].extend(values)
                _merge_statistics(ch, stats[c])
                ch.current_state = states[c]
                ch.temperature = float(temps[c])
            self._saved_from = len(batch.saved)
            return self.chains

    class CudaVanillaSampler(_CudaSampler):
        """``VanillaSampler`` (samplers/_samplers.py:242-276) on the device."""

        def run(self, n_iterations, burnin_iterations=0, save_every=100, verbose=True, print_every=100):
            self._open()
            self.batch.advance(n_iterations, burnin_iterations, save_every)
            return self._close()

    class CudaParallelTempering(_CudaSampler):
        """``ParallelTempering`` (samplers/_samplers.py:279-365) on the device: the reference's ladder rule, its
        window loop (quirk Q3 included) and its swap round of ``n_chains`` sequential random-pair attempts, replayed
        by one kernel (``bbb_pt_swap``); temperatures are exchanged, states never move."""

        def __init__(self, temperature_max=5, chains_with_unit_temperature=0.4, swap_every=500, device=device, seed=seed):
            super().__init__(device=device, seed=seed)
            self._temperature_max = temperature_max
            self._chains_with_unit_tempeature = chains_with_unit_temperature
            self._swap_every = swap_every

        def on_initialize(self, chains):  # reference :315-326
            n_chains = len(chains)
            ones = np.ones(max(2, int(n_chains * self._chains_with_unit_tempeature)) - 1)
            temperatures = np.concatenate((ones, np.geomspace(1, self._temperature_max, n_chains - len(ones))))
            for i, chain in enumerate(chains):
                chain.temperature = float(temperatures[i])

        def run(self, n_iterations, burnin_iterations=0, save_every=100, verbose=True, print_every=100):
            from ._engine import pt_swap

            self._open()
            batch = self.batch
            start = batch.iteration
            while True:  # literal loop of the reference (:344-365)
                iteration = batch.iteration - start
                n_it = min(self._swap_every, n_iterations - iteration)
                burnin_it = max(0, min(self._swap_every, burnin_iterations - iteration))
                if n_it > 0:
                    batch.advance(n_it, burnin_it, save_every)
                t_out, _ = pt_swap(batch.misfit_logdet_T(), batch.seed_global, batch.swap_round, 0, batch.n_chains)
                batch.engine.temperature.copy_(t_out)
                batch.swap_round += 1
                if iteration >= n_iterations:
                    break
            return self._close()

    return CudaVanillaSampler, CudaParallelTempering
