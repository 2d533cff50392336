"""``ParameterSpace`` and ``Parameterization`` descriptors (reference
src/bayesbay/parameterization/_parameter_space.py, _parameterization.py)."""
from __future__ import annotations

from typing import Dict, List, Union

import numpy as np

from ._state import ParameterSpaceState, State
from .exceptions import DeviceUnsupported
from .perturbations import BirthPerturbation, DeathPerturbation, ParamPerturbation, ParamSpacePerturbation
from .prior import Prior

__all__ = ["ParameterSpace", "Parameterization"]


class ParameterSpace:
    """A (possibly trans-dimensional) vector of named parameters.

    ``ParameterSpace(name, n_dimensions=None, n_dimensions_min=1, n_dimensions_max=10,
    n_dimensions_init_range=0.3, parameters=None)`` -- reference
    parameterization/_parameter_space.py:16-371.  Cells are padded to ``n_dimensions_max`` on the
    device.  Nested parameter spaces are not supported (ragged recursive state)."""

    def __init__(self, name: str, n_dimensions: int = None, n_dimensions_min: int = 1, n_dimensions_max: int = 10,
                 n_dimensions_init_range: float = 0.3, parameters: List[Prior] = None):
        self._name = name
        self._trans_d = n_dimensions is None
        self._n_dimensions = n_dimensions
        self._n_dimensions_min = n_dimensions_min if self._trans_d else n_dimensions
        self._n_dimensions_max = n_dimensions_max if self._trans_d else n_dimensions
        self._n_dimensions_init_range = n_dimensions_init_range
        self._parameters: Dict[str, Prior] = {}
        for p in parameters or []:
            if isinstance(p, ParameterSpace):
                raise DeviceUnsupported("nested parameter spaces are not supported by the device engine")
            self._parameters[p.name] = p
        self._is_leaf = True
        self._init_perturbation_funcs()
        self._repr_args = {"name": name, "parameters": list(self._parameters.values())}
        if self._trans_d:
            self._repr_args.update(n_dimensions_min=n_dimensions_min, n_dimensions_max=n_dimensions_max,
                                   n_dimensions_init_range=n_dimensions_init_range)
        else:
            self._repr_args["n_dimensions"] = n_dimensions

    name = property(lambda self: self._name)
    trans_d = property(lambda self: self._trans_d)
    is_leaf = property(lambda self: self._is_leaf)
    parameters = property(lambda self: self._parameters)
    perturbation_funcs = property(lambda self: self._perturbation_funcs)
    perturbation_weights = property(lambda self: self._perturbation_weights)

    def _inner_perturbations(self):
        funcs, weights = [], []
        if self.trans_d:
            funcs += [BirthPerturbation(self), DeathPerturbation(self)]
            weights += [1, 1]
        if self.parameters:
            funcs.append(ParamPerturbation(self.name, list(self.parameters.values())))
            weights.append(3)
        return funcs, weights

    def _init_perturbation_funcs(self):
        # weights birth 1, death 1, values 3 (reference _parameter_space.py:324-352)
        funcs, weights = self._inner_perturbations()
        self._perturbation_funcs = [ParamSpacePerturbation(self.name, funcs, weights)]
        self._perturbation_weights = [sum(weights)]

    def initialize(self, position=None) -> Union[ParameterSpaceState, List[ParameterSpaceState]]:
        """Draw a starting state on the device (reference ``_initialize``, :146-160)."""
        from ._batch import sample_space_states

        if position is None:
            return sample_space_states(self, 1)[0]
        return sample_space_states(self, len(position))

    def log_prior(self, param_space_state: ParameterSpaceState) -> float:
        """Joint log prior of the space's parameters (reference :189-195), evaluated by the CUDA
        log-prior kernel."""
        from ._batch import device_log_prior

        return device_log_prior(self, [param_space_state])[0]

    def birth(self, ps_state):
        raise TypeError("birth proposals run inside the CUDA engine (csrc/propose.cuh)")

    death = birth

    def __repr__(self) -> str:
        args = ", ".join(f"{k}={v}" for k, v in self._repr_args.items() if k != "name")
        return f"{self._repr_args['name']}({args})"


class Parameterization:
    """Collection of parameter spaces (reference parameterization/_parameterization.py:11-84)."""

    def __init__(self, parameter_space: Union[ParameterSpace, List[ParameterSpace]]):
        spaces = parameter_space if isinstance(parameter_space, list) else [parameter_space]
        names = [ps.name for ps in spaces]
        if len(names) != len(set(names)):
            raise ValueError("duplicate parameter space names found")
        self.parameter_spaces = {ps.name: ps for ps in spaces}
        self._perturbation_funcs, self._perturbation_weights = [], []
        for ps in spaces:
            self._perturbation_funcs.extend(ps.perturbation_funcs)
            self._perturbation_weights.extend(ps.perturbation_weights)

    perturbation_funcs = property(lambda self: self._perturbation_funcs)
    perturbation_weights = property(lambda self: self._perturbation_weights)

    def initialize(self) -> State:
        return State({name: ps.initialize() for name, ps in self.parameter_spaces.items()})

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(parameter_spaces={list(self.parameter_spaces.values())})"
