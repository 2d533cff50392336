"""Samplers (reference src/bayesbay/samplers/_samplers.py) driving the CUDA engine.

``Sampler`` keeps the reference's contract -- ``initialize(chains)``, the ``parallel_config``
property, ``run(...) -> List[BaseMarkovChain]`` called by ``BaseBayesianInversion.run``
(src/bayesbay/_bayes_inversion.py:236-249) -- but ``advance_chain`` advances ALL chains in lock step
on the device instead of mapping ``MarkovChain.advance_chain`` over a joblib/loky process pool
(samplers/_samplers.py:191-239).

Callback semantics: the per-chain-per-iteration hooks ``on_begin_iteration`` / ``on_end_iteration``
cannot survive batching; the built-in samplers' behaviour is reproduced natively (Vanilla: no-ops,
ParallelTempering: swap round, SimulatedAnnealing: temperature schedule) and registering user
callbacks with ``add_on_begin_iteration`` / ``add_on_end_iteration`` raises.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from typing import Callable, List

import numpy as np

__all__ = ["Sampler", "VanillaSampler", "ParallelTempering", "SimulatedAnnealing"]


class Sampler(ABC):
    def __init__(self):
        self._extra_on_initialize = []
        self._extra_on_end_advance_chain = []
        self._parallel_config = {}

    # ---- called by BaseBayesianInversion.run ------------------------------
    def initialize(self, chains):
        self._chains = chains
        self._owner = getattr(chains[0], "_batch_owner", None) if chains else None
        if self._owner is None:
            raise TypeError("these chains are not backed by a device batch; build them with bayesbay_b200.BayesianInversion")
        self.on_initialize(chains)
        for func in self._extra_on_initialize:
            func(self, chains)

    def add_on_initialize(self, func: Callable):
        self._extra_on_initialize.append(func)

    def add_on_begin_iteration(self, func: Callable):
        raise TypeError("per-iteration Python callbacks cannot run inside the batched CUDA step")

    add_on_end_iteration = add_on_begin_iteration

    def add_on_end_advance_chain(self, func: Callable):
        # quirk Q4: the reference never invokes these (advance_chain calls on_end_advance_chain directly)
        self._extra_on_end_advance_chain.append(func)

    @abstractmethod
    def on_initialize(self, chains):
        raise NotImplementedError

    @abstractmethod
    def on_end_advance_chain(self):
        raise NotImplementedError

    @abstractmethod
    def run(self, n_iterations, burnin_iterations=0, save_every=100, verbose=True, print_every=100):
        raise NotImplementedError

    @property
    def chains(self):
        return self._chains

    @property
    def parallel_config(self) -> dict:
        return self._parallel_config

    @parallel_config.setter
    def parallel_config(self, parallel_config: dict):
        assert isinstance(parallel_config, dict)
        self._parallel_config = parallel_config

    def advance_chain(self, n_iterations, burnin_iterations=0, save_every=100, verbose=True, print_every=100):
        """Advance every chain ``n_iterations`` steps on the device, then ``on_end_advance_chain()``
        (reference samplers/_samplers.py:191-239)."""
        owner = self._owner
        done = 0
        while done < n_iterations:
            n = n_iterations - done
            if verbose:
                n = min(n, print_every - owner.batch.iteration % print_every)
            owner.batch.advance(n, burnin_iterations, save_every)
            done += n
            if verbose and owner.batch.iteration % print_every == 0:
                owner._mark_dirty()
                owner._sync_chains(states=False)
                for ch in self._chains:
                    ch.print_statistics()
        self.on_end_advance_chain()
        owner._mark_dirty()
        return self._chains


class VanillaSampler(Sampler):
    """High-level class for the Vanilla MCMC sampler (reference :242-276)."""

    def on_initialize(self, chains):
        self._owner.batch.resumed = False

    def on_end_advance_chain(self):
        pass

    def run(self, n_iterations, burnin_iterations=0, save_every=100, verbose=True, print_every=100):
        return self.advance_chain(n_iterations=n_iterations, burnin_iterations=burnin_iterations,
                                  save_every=save_every, verbose=verbose, print_every=print_every)


class ParallelTempering(Sampler):
    """Parallel tempering (reference :279-365).

    ``temperatures`` (extension, optional): explicit per-chain ladder instead of the reference's
    ``on_initialize`` rule (quirk Q5) -- needed e.g. for "16 temperatures x 512 chains".
    The swap round gathers ``(misfit, logdet, T)`` of all chains (NCCL all-gather when chains are
    sharded over ranks) and replays the reference's ``n_chains`` sequential random-pair attempts in
    one kernel; temperatures are swapped, states never move."""

    def __init__(self, temperature_max=5, chains_with_unit_temperature=0.4, swap_every=500, temperatures=None):
        super().__init__()
        self._temperature_max = temperature_max
        self._chains_with_unit_tempeature = chains_with_unit_temperature
        self._swap_every = swap_every
        self._temperatures = None if temperatures is None else np.asarray(temperatures, dtype=np.float64)
        self._round = 0
        self.n_swaps_accepted = 0

    def on_initialize(self, chains):
        owner = self._owner
        if owner.batch.resumed:
            # continuing a checkpoint: the per-chain temperatures ARE the accumulated swap permutation
            owner.batch.resumed = False
            for c, t in enumerate(owner.batch.temperatures()):
                chains[c].temperature = float(t)
            return
        n_total = owner.n_chains_global
        if self._temperatures is not None:
            if self._temperatures.size != n_total:
                raise ValueError(f"`temperatures` has {self._temperatures.size} entries for {n_total} chains")
            temps = self._temperatures
        else:
            # reference on_initialize (:315-326) incl. its failure for a single chain
            ones = np.ones(max(2, int(n_total * self._chains_with_unit_tempeature)) - 1)
            if n_total - ones.size <= 0:
                raise ValueError("ParallelTempering needs at least 2 chains")
            temps = np.concatenate((ones, np.geomspace(1, self._temperature_max, n_total - ones.size)))
        local = temps[owner.chain_id0: owner.chain_id0 + len(chains)]
        owner.batch.set_temperatures(local)
        for c, chain in enumerate(chains):
            chain.temperature = float(local[c])

    def on_end_advance_chain(self):
        owner = self._owner
        local = owner.batch.misfit_logdet_T()
        gathered = owner.all_gather(local)
        from ._engine import pt_swap

        # the round number (Philox counter of the swap stream) lives with the chains so that it is checkpointed
        _, n_acc = pt_swap(gathered, owner.batch.seed_global, owner.batch.swap_round, owner.chain_id0,
                           owner.batch.n_chains, out=owner.batch.engine.temperature)
        owner.batch.swap_round += 1
        self._round += 1
        self._last_n_acc = n_acc

    def run(self, n_iterations, burnin_iterations=0, save_every=100, verbose=True, print_every=100):
        # literal loop of the reference (:344-365), quirk Q3 included: the burn-in handed to each window is
        # window-relative while the chains count iterations globally, and one zero-length advance + swap
        # round runs after the last window.
        owner = self._owner
        while True:
            iteration = owner.batch.iteration
            n_it = min(self._swap_every, n_iterations - iteration)
            burnin_it = max(0, min(self._swap_every, burnin_iterations - iteration))
            self.advance_chain(n_iterations=n_it, burnin_iterations=burnin_it, save_every=save_every,
                               verbose=verbose, print_every=print_every)
            if iteration >= n_iterations:
                break
        return self.chains


class SimulatedAnnealing(Sampler):
    """Simulated annealing (reference :368-448): exponential cooling from ``temperature_start`` to 1
    over the first ``cooling_fraction`` of the burn-in, evaluated per iteration on the device."""

    def __init__(self, temperature_start=10, cooling_fraction=0.8):
        super().__init__()
        self.temperature_start = temperature_start
        self.cooling_fraction = float(cooling_fraction)

    def on_initialize(self, chains):
        if self._owner.batch.resumed:  # continuing a checkpoint: the schedule is a function of the iteration counter
            self._owner.batch.resumed = False
            return
        self._owner.batch.set_temperatures(np.full(len(chains), float(self.temperature_start)))
        for chain in chains:
            chain.temperature = self.temperature_start

    def on_end_advance_chain(self):
        pass

    def run(self, n_iterations, burnin_iterations=0, save_every=100, verbose=True, print_every=100):
        self.burnin_iterations = burnin_iterations
        self.cooling_iterations = 0
        self.cooling_rate = 0.0
        if burnin_iterations != 0:
            cf = min(max(self.cooling_fraction, 0.0), 1.0)
            self.cooling_iterations = max(1, min(burnin_iterations, int(round(cf * burnin_iterations))))
            self.cooling_rate = math.log(self.temperature_start) / self.cooling_iterations
        self._owner.batch.engine.set_annealing(self.temperature_start, self.cooling_rate, self.cooling_iterations,
                                               burnin_iterations)
        try:
            self.advance_chain(n_iterations=n_iterations, burnin_iterations=burnin_iterations,
                               save_every=save_every, verbose=verbose, print_every=print_every)
        finally:
            self._owner.batch.engine.set_annealing(0.0, 0.0, 0, 0)
        return self.chains
