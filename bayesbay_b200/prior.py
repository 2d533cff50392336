"""Prior descriptors with the reference's names and signatures (src/bayesbay/prior/_prior.py).

The objects carry hyper-parameters to the device engine; sampling, random-walk perturbation and
the prior ratios of the MCMC loop run in CUDA (``csrc/common.cuh``, ``csrc/propose.cuh``).  Only the
closed-form ``log_prior(value)`` of a single scalar is evaluated on the host, as in the reference's
interactive use (tests/22_log_like_func.py:24-31); the batched joint log prior of a parameter space
runs on the device (``ParameterSpace.log_prior`` -> ``bbb_logprior``).
"""
from __future__ import annotations

import math
from numbers import Number

import numpy as np

from .exceptions import DeviceUnsupported

__all__ = ["Prior", "UniformPrior", "GaussianPrior", "LaplacePrior", "CustomPrior",
           "UniformParameter", "GaussianParameter", "CustomParameter"]


def _interp(xp, x, y):
    """``interpolate_linear_1d`` (reference _utils_1d.pyx:68-84): clamp outside, linear inside."""
    if xp <= x[0]:
        return float(y[0])
    if xp >= x[-1]:
        return float(y[-1])
    i = int(np.searchsorted(x, xp, side="left")) - 1
    i = max(i, 0)
    return float(y[i] + (xp - x[i]) * (y[i + 1] - y[i]) / (x[i + 1] - x[i]))


class Prior:
    """Base class (reference prior/_prior.py:24-297).

    Hyper-parameters are scalars, or -- with ``position`` (1-D, ascending) -- arrays of ``len(position)``
    that are linearly interpolated at the position of the cell (reference ``get_hyper_param`` :257-297)."""

    kind = None
    _hyper_names = ()

    def __init__(self, name, perturb_std, perturb_std_birth=None, position=None, **hyper):
        self._name = name
        if position is not None:
            self.position = np.array(position, dtype=float)
            if self.position.ndim != 1:
                raise DeviceUnsupported("only 1-D `position` arrays (Voronoi1D) are supported by the device engine")
            if self.position.size < 2 or np.any(np.diff(self.position) <= 0):
                raise ValueError("`position` should be strictly ascending with at least two entries")
        else:
            self.position = None
        hyper = dict(hyper, perturb_std=perturb_std,
                     perturb_std_birth=perturb_std if perturb_std_birth is None else perturb_std_birth)
        for key, val in hyper.items():
            if np.isscalar(val) and isinstance(val, Number):
                setattr(self, key, float(val))
            else:
                assert self.position is not None, f"`{key}` should be a scalar when `position` is `None`"
                arr = np.array(val, dtype=float)
                assert arr.shape == self.position.shape, (
                    f"`{key}` should either be a scalar or have the same length as `position`")
                setattr(self, key, arr)
        self._hyper_names = tuple(hyper)
        if np.any(np.asarray(self.perturb_std) <= 0) or np.any(np.asarray(self.perturb_std_birth) <= 0):
            raise ValueError("perturb_std must be positive")
        self._repr_args = dict(name=name, **{k: v for k, v in hyper.items()}, position=position)

    def get_hyper_param(self, hyper_param: str, position=None):
        val = getattr(self, hyper_param)
        if np.isscalar(val):
            return val
        if position is None:
            raise ValueError(f"`{hyper_param}` is dependent on position but got None for it")
        return _interp(float(position), self.position, val)

    def has_hyper_param(self, hyper_param: str) -> bool:
        return hasattr(self, hyper_param)

    @property
    def name(self) -> str:
        return self._name

    def get_perturb_std(self, position=None):
        return self.get_hyper_param("perturb_std", position)

    def get_perturb_std_birth(self, position=None):
        return self.get_hyper_param("perturb_std_birth", position)

    def _table(self, a_name, b_name):
        """(position, a, b, perturb_std, perturb_std_birth) arrays for the device, or None if all scalar."""
        if self.position is None:
            return None
        col = lambda v: np.broadcast_to(np.asarray(v, dtype=float), self.position.shape).copy()  # noqa: E731
        return dict(position=self.position.copy(), a=col(getattr(self, a_name)), b=col(getattr(self, b_name)),
                    perturb_std=col(self.perturb_std), perturb_std_birth=col(self.perturb_std_birth))

    def log_prior(self, value, position=None):
        raise NotImplementedError

    def _device_params(self):
        """(kind, a, b) for the device prior table."""
        raise NotImplementedError

    def __repr__(self):
        args = ", ".join(f"{k}={v}" for k, v in self._repr_args.items())
        return f"{self.__class__.__name__}({args})"


class UniformPrior(Prior):
    """``UniformPrior(name, vmin, vmax, perturb_std, perturb_std_birth=None, position=None)``
    (reference prior/_prior.py:300-481)."""

    kind = "uniform"

    def __init__(self, name, vmin, vmax, perturb_std, perturb_std_birth=None, position=None):
        super().__init__(name, perturb_std, perturb_std_birth, position, vmin=vmin, vmax=vmax)
        if np.any(np.asarray(self.vmin) >= np.asarray(self.vmax)):
            raise ValueError("vmin should be less than vmax")
        self.delta = np.asarray(self.vmax) - np.asarray(self.vmin) if self.position is not None else self.vmax - self.vmin

    def get_vmin_vmax(self, position=None):
        return self.get_hyper_param("vmin", position), self.get_hyper_param("vmax", position)

    def get_delta(self, position=None):
        vmin, vmax = self.get_vmin_vmax(position)
        return vmax - vmin

    def log_prior(self, value, position=None):
        vmin, vmax = self.get_vmin_vmax(position)
        return -math.log(vmax - vmin) if vmin <= value <= vmax else -math.inf

    def _device_params(self):
        return "uniform", self.vmin, self.vmax, self._table("vmin", "vmax")


class GaussianPrior(Prior):
    """``GaussianPrior(name, mean, std, perturb_std, perturb_std_birth=None, position=None)``
    (reference prior/_prior.py:484-682)."""

    kind = "gaussian"

    def __init__(self, name, mean, std, perturb_std, perturb_std_birth=None, position=None):
        super().__init__(name, perturb_std, perturb_std_birth, position, mean=mean, std=std)
        if np.any(np.asarray(self.std) <= 0):
            raise ValueError("std must be positive")

    def get_mean(self, position=None):
        return self.get_hyper_param("mean", position)

    def get_std(self, position=None):
        return self.get_hyper_param("std", position)

    def log_prior(self, value, position=None):
        mean, std = self.get_mean(position), self.get_std(position)
        return -0.5 * math.log(2 * math.pi) - math.log(std) - 0.5 * ((value - mean) / std) ** 2

    def _device_params(self):
        return "gaussian", self.mean, self.std, self._table("mean", "std")


class LaplacePrior(Prior):
    """``LaplacePrior(name, mean, scale, perturb_std, perturb_std_birth=None, position=None)``
    (reference prior/_prior.py:685-881)."""

    kind = "laplace"

    def __init__(self, name, mean, scale, perturb_std, perturb_std_birth=None, position=None):
        super().__init__(name, perturb_std, perturb_std_birth, position, mean=mean, scale=scale)
        if np.any(np.asarray(self.scale) <= 0):
            raise ValueError("scale must be positive")

    def get_mean(self, position=None):
        return self.get_hyper_param("mean", position)

    def get_scale(self, position=None):
        return self.get_hyper_param("scale", position)

    def log_prior(self, value, position=None):
        mean, scale = self.get_mean(position), self.get_scale(position)
        return -math.log(2) - math.log(scale) - abs(value - mean) / scale

    def _device_params(self):
        return "laplace", self.mean, self.scale, self._table("mean", "scale")


class CustomPrior(Prior):
    """``CustomPrior(name, log_prior, sample, perturb_std, ...)`` (reference prior/_prior.py:884-1014).

    Arbitrary Python callables cannot run inside a CUDA kernel and this package has no CPU fallback:
    the object can be built (API compatibility) but using it in an inversion raises at compile time."""

    kind = "custom"

    def __init__(self, name, log_prior, sample, perturb_std, perturb_std_birth=None, position=None):
        super().__init__(name, perturb_std, perturb_std_birth, position)
        self._log_prior = log_prior
        self._sample = sample

    def log_prior(self, value, position=None):
        try:
            return self._log_prior(value, position)
        except TypeError:
            return self._log_prior(value)

    def _device_params(self):
        raise DeviceUnsupported(
            f"CustomPrior {self.name!r}: user callables cannot run on the device; use Uniform/Gaussian/LaplacePrior")


# pre-0.2 names of the prior classes (reference CHANGELOG.md:91-94)
UniformParameter = UniformPrior
GaussianParameter = GaussianPrior
CustomParameter = CustomPrior
