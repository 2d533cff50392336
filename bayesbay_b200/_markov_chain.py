"""``MarkovChain`` / ``BaseMarkovChain`` views (reference src/bayesbay/_markov_chain.py).

In the reference a chain object *is* the sampler loop; here the loop runs on the device for all
chains at once (``_batch.ChainBatch``) and a chain object is a per-chain view exposing what the
facade and users read: ``saved_states``, ``statistics``, ``current_state``, ``temperature``,
``ith_iteration``, ``print_statistics()``."""
from __future__ import annotations

from collections import defaultdict
from typing import Union

__all__ = ["BaseMarkovChain", "MarkovChain"]


class BaseMarkovChain:
    def __init__(self, id: Union[int, str], batch=None, index: int = 0, parameterization=None, log_likelihood=None,
                 perturbation_funcs=None, perturbation_weights=None, temperature: float = 1, save_dpred: bool = True):
        self.id = id
        self._batch = batch
        self._index = index
        self.parameterization = parameterization
        self.log_likelihood = log_likelihood
        self.save_dpred = save_dpred
        self._temperature = temperature
        self.set_perturbation_funcs(perturbation_funcs or [], perturbation_weights or [])
        self._saved_states = defaultdict(list)
        self._statistics = {
            "n_proposed_models": defaultdict(lambda: defaultdict(float)),
            "n_accepted_models": defaultdict(lambda: defaultdict(float)),
            "n_proposed_models_total": 0,
            "n_accepted_models_total": 0,
            "exceptions": defaultdict(int),
        }
        self._current_state = None

    def set_perturbation_funcs(self, funcs, weights):
        self.perturbation_funcs = funcs
        self.perturbation_types = [f.__name__ for f in funcs]
        self.perturbation_weights = weights

    def _ensure_synced(self, states=False):
        owner = getattr(self, "_batch_owner", None)
        if owner is not None and (owner._dirty or (states and self._current_state is None)):
            owner._sync_chains(states=states)

    @property
    def current_state(self):
        self._ensure_synced(states=True)
        return self._current_state

    @current_state.setter
    def current_state(self, state):
        self._current_state = state

    @property
    def temperature(self):
        """The current temperature used in the log likelihood ratio"""
        self._ensure_synced()
        return self._temperature

    @temperature.setter
    def temperature(self, value):
        self._temperature = value
        if self._current_state is not None:
            self._current_state.temperature = value

    @property
    def saved_states(self):
        """All the states saved so far"""
        self._ensure_synced()
        return self._saved_states

    @property
    def statistics(self):
        self._ensure_synced()
        return self._statistics

    @property
    def ith_iteration(self):
        return self.statistics["n_proposed_models_total"]

    def print_statistics(self):
        """Same report as the reference (src/bayesbay/_markov_chain.py:159-197)."""
        st = self.statistics
        tot_a, tot_p = st["n_accepted_models_total"], st["n_proposed_models_total"]
        print(f"Chain ID: {self.id}\nTEMPERATURE: {self.temperature}\nEXPLORED MODELS: {tot_p}")
        print("ACCEPTANCE RATE: %d/%d (%.2f %%)" % (tot_a, tot_p, tot_a / max(tot_p, 1) * 100))
        print("PARTIAL ACCEPTANCE RATES:")
        for ptype in sorted(st["n_proposed_models"]):
            p, a = st["n_proposed_models"][ptype], st["n_accepted_models"][ptype]
            if isinstance(p, dict):
                print(f"\t{ptype}:")
                for sub in sorted(p):
                    print(f"\t\t{sub}: {a[sub]:.2f}/{p[sub]:.2f} ({a[sub] / p[sub] * 100:.2f}%)")
            else:
                print("\t%s: %d/%d (%.2f%%)" % (ptype, a, p, a / p * 100))

    def advance_chain(self, *args, **kwargs):
        raise TypeError("chains advance together on the device: use BayesianInversion.run() / Sampler.advance_chain()")

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(id={self.id}, temperature={self.temperature}, "
                f"n_proposed_models_total={self.statistics['n_proposed_models_total']}, "
                f"n_accepted_models_total={self.statistics['n_accepted_models_total']})")


class MarkovChain(BaseMarkovChain):
    """High-level chain view (reference src/bayesbay/_markov_chain.py:314-452)."""
