"""Host-visible chain state containers.

Same names, fields and dict shapes as the reference's ``State`` / ``ParameterSpaceState`` /
``DataNoiseState`` (src/bayesbay/_state.py:11-461): these are what ``chain.current_state``,
``walkers_starting_states`` and the saved samples are made of.  On the device the same information
lives in padded structure-of-arrays buffers (see ``_engine.py``); these classes are only the
host-side exchange format.
"""
from __future__ import annotations

from collections import namedtuple
from numbers import Number
from typing import Any, Dict, List, Union

import numpy as np


class DataNoiseState(namedtuple("DataNoiseState", ["std", "correlation"])):
    """Noise standard deviation (and correlation) of one target (reference _state.py:11-46)."""

    __slots__ = ()

    def __new__(cls, std, correlation=None):
        return super().__new__(cls, std, correlation)

    def copy(self) -> "DataNoiseState":
        return DataNoiseState(self.std, self.correlation)

    def todict(self, name: str) -> dict:
        out = {f"{name}.std": self.std}
        if self.correlation is not None:
            out[f"{name}.correlation"] = self.correlation
        return out


class ParameterSpaceState:
    """Values of one parameter space: ``n_dimensions`` cells, one array per parameter, plus
    ``"discretization"`` for Voronoi spaces (reference _state.py:50-232)."""

    def __init__(self, n_dimensions: int, param_values: Dict[str, Any] = None, cache: Dict[str, Any] = None):
        if not isinstance(n_dimensions, (int, np.integer)):
            raise TypeError("n_dimensions should be an int")
        if param_values is None:
            param_values = {}
        if not isinstance(param_values, dict):
            raise TypeError("param_values should be a dict")
        self.n_dimensions = int(n_dimensions)
        self.param_values: Dict[str, Any] = {}
        self.cache: Dict[str, Any] = {} if cache is None else cache
        for name, values in param_values.items():
            if len(values) != self.n_dimensions:
                raise ValueError(
                    f"parameter {name} should have the same length as `n_dimensions` "
                    f"({self.n_dimensions}) but have {len(values)} instead"
                )
            self.set_param_values(name, values)

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return self.get_param_values(idx)
        if isinstance(idx, list):
            return ParameterSpaceState(len(idx), {k: v[idx] for k, v in self.param_values.items()})
        raise TypeError("`idx` should be either a string or a list")

    def set_param_values(self, param_name: str, values):
        if not isinstance(param_name, str):
            raise ValueError("`param_name` should be a string")
        if not isinstance(values, np.ndarray):
            raise TypeError("parameter values should be a numpy ndarray (nested parameter spaces are not "
                            "supported by the device engine)")
        self.param_values[param_name] = values
        setattr(self, param_name, values)

    def get_param_values(self, param_name: str):
        if not isinstance(param_name, str):
            raise ValueError("`param_name` should be a string")
        return self.param_values.get(param_name, None)

    def saved_in_cache(self, name: str) -> bool:
        return name in self.cache

    def load_from_cache(self, name: str):
        return self.cache[name]

    def save_to_cache(self, name: str, value):
        self.cache[name] = value

    def copy(self) -> "ParameterSpaceState":
        new = ParameterSpaceState(self.n_dimensions, {k: v.copy() for k, v in self.param_values.items()})
        new.cache = self.cache.copy()
        return new

    def todict(self, name: str) -> dict:
        out = {f"{name}.n_dimensions": self.n_dimensions}
        for k, v in self.param_values.items():
            out[f"{name}.{k}"] = v
        return out

    def __repr__(self):
        return f"ParameterSpaceState(n_dimensions={self.n_dimensions}, param_values={self.param_values})"

    def __eq__(self, other):
        return (isinstance(other, ParameterSpaceState) and self.n_dimensions == other.n_dimensions
                and self.param_values.keys() == other.param_values.keys()
                and all(np.array_equal(v, other.param_values[k]) for k, v in self.param_values.items()))


class State:
    """Full chain state: parameter-space states + data-noise states, temperature, cache (holds
    ``"{target}.dpred"``) and extra storage (reference _state.py:236-461)."""

    def __init__(self, param_values: Dict[str, Union[ParameterSpaceState, DataNoiseState]] = None,
                 temperature: float = 1, cache: Dict[str, Any] = None, extra_storage: Dict[str, Any] = None):
        if not isinstance(temperature, Number):
            raise TypeError("temperature should be a number")
        if param_values is None:
            param_values = {}
        if not isinstance(param_values, dict):
            raise TypeError("param_values should be a dict")
        self.param_values: Dict[str, Any] = {}
        self.temperature = temperature
        self.cache = {} if cache is None else cache
        self.extra_storage = {} if extra_storage is None else extra_storage
        for name, values in param_values.items():
            self.set_param_values(name, values)

    def __getitem__(self, name: str):
        return self.get_param_values(name)

    def set_param_values(self, param_name: str, values):
        if not isinstance(param_name, str):
            raise ValueError("`param_name` should be a string")
        if not isinstance(values, (ParameterSpaceState, DataNoiseState)):
            raise TypeError("parameter values should either be a ParameterSpaceState or a `DataNoiseState` instance")
        self.param_values[param_name] = values
        setattr(self, param_name, values)

    def get_param_values(self, param_name: str):
        if not isinstance(param_name, str):
            raise ValueError("`param_name` should be a string")
        return self.param_values.get(param_name, None)

    def saved_in_cache(self, name: str) -> bool:
        return name in self.cache

    def load_from_cache(self, name: str):
        return self.cache[name]

    def save_to_cache(self, name: str, value):
        self.cache[name] = value

    def saved_in_extra_storage(self, name: str) -> bool:
        return name in self.extra_storage

    def load_from_extra_storage(self, name: str):
        return self.extra_storage[name]

    def save_to_extra_storage(self, name: str, value):
        self.extra_storage[name] = value

    def _vars(self) -> dict:
        out = {}
        for k, v in self.param_values.items():
            out.update(v.todict(k))
        out.update(self.extra_storage)
        return out

    def __iter__(self):
        return iter(self._vars())

    def items(self):
        return self._vars().items()

    def copy(self, keep_dpred: bool = False) -> "State":
        new = State({k: v.copy() for k, v in self.param_values.items()}, temperature=self.temperature)
        if keep_dpred:
            new.cache = {k: v for k, v in self.cache.items() if "dpred" in k}
        return new

    def __repr__(self):
        return f"State(param_values={self.param_values}, temperature={self.temperature})"
