"""Perturbation descriptors (reference src/bayesbay/perturbations/).

The proposals themselves run on the device (``csrc/propose.cuh``); these objects only carry the
structure (which moves exist, their weights) and the ``__name__`` strings that key the per-chain
statistics dictionaries (``_save_statistics``, src/bayesbay/_markov_chain.py:136-157)."""
from __future__ import annotations

from typing import List

__all__ = ["Perturbation", "ParamPerturbation", "BirthPerturbation", "DeathPerturbation", "NoisePerturbation",
           "ParamSpacePerturbation"]


class Perturbation:
    """Base class (perturbations/_base_perturbation.py:7-90)."""

    move_kind = None  # device inner-move id (birth 0, death 1, values 2, sites 3)

    @property
    def type(self) -> str:
        return self.__class__.__name__

    @property
    def __name__(self) -> str:
        return self.type

    def perturb(self, state):
        raise TypeError(f"{self.__name__} runs inside the CUDA engine (csrc/propose.cuh); there is no host implementation")

    def __call__(self, state):
        return self.perturb(state)

    def __repr__(self):
        return self.__name__


class ParamPerturbation(Perturbation):
    """Value move of every listed parameter at one random cell, or the site move when the listed
    "parameter" is the discretization itself (perturbations/_param_values.py:10-99)."""

    def __init__(self, param_space_name: str, parameters: list):
        self.param_space_name = param_space_name
        self.parameters = parameters
        self.move_kind = 3 if any(hasattr(p, "birth_from") for p in parameters) else 2

    @property
    def __name__(self) -> str:
        names = [f"{p.name}.discretization" if hasattr(p, "trans_d") else f"{self.param_space_name}.{p.name}"
                 for p in self.parameters]
        return f"{self.type}({str(names) if len(names) != 1 else str(names[0])})"


class BirthPerturbation(Perturbation):
    move_kind = 0

    def __init__(self, parameter_space):
        self.param_space = parameter_space
        self.param_space_name = parameter_space.name

    @property
    def __name__(self) -> str:
        return f"{self.type}({self.param_space_name})"


class DeathPerturbation(Perturbation):
    move_kind = 1

    def __init__(self, parameter_space):
        self.param_space = parameter_space
        self.param_space_name = parameter_space.name

    @property
    def __name__(self) -> str:
        return f"{self.type}({self.param_space_name})"


class NoisePerturbation(Perturbation):
    """Random walk of the noise level of every hierarchical target (perturbations/_data_noise.py:10-86)."""

    def __init__(self, targets: list):
        self.targets = targets

    @property
    def __name__(self) -> str:
        names = [t.name for t in self.targets]
        return f"{self.type}({str(names) if len(names) > 1 else str(names[0])})"


class ParamSpacePerturbation(Perturbation):
    """One weighted choice among the inner perturbations of a parameter space
    (perturbations/_param_space.py:10-112)."""

    def __init__(self, param_space_name: str, perturbations: List[Perturbation], perturbation_weights: list):
        assert len(perturbations) == len(perturbation_weights), (
            "The number of perturbations and perturbation weights must be equal")
        self.param_space_name = param_space_name
        self._perturbations_functions = perturbations
        self._perturbation_weights = perturbation_weights

    @property
    def perturbation_funcs(self):
        return self._perturbations_functions

    @perturbation_funcs.setter
    def perturbation_funcs(self, value):
        self._perturbations_functions = value

    @property
    def perturbation_weights(self):
        return self._perturbation_weights

    @perturbation_weights.setter
    def perturbation_weights(self, value):
        self._perturbation_weights = value

    @property
    def __name__(self) -> str:
        return f"ParamSpacePerturbation(param_space_name={self.param_space_name})"
