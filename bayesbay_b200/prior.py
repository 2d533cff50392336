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


def _scalar(name, value):
    if value is None or not np.isscalar(value) or not isinstance(value, Number):
        raise DeviceUnsupported(
            f"`{name}` must be a scalar: position-dependent hyper-parameters (prior/_prior.py:257-297) are not "
            "implemented by the device engine yet (SURVEY.md section 8f rank 3)")
    return float(value)


class Prior:
    """Base class (reference prior/_prior.py:24-297)."""

    kind = None

    def __init__(self, name, perturb_std, perturb_std_birth=None, position=None, **hyper):
        if position is not None:
            raise DeviceUnsupported("position-dependent priors are not implemented by the device engine yet")
        self._name = name
        self.position = None
        self.perturb_std = _scalar("perturb_std", perturb_std)
        self.perturb_std_birth = self.perturb_std if perturb_std_birth is None else _scalar("perturb_std_birth", perturb_std_birth)
        if not self.perturb_std > 0 or not self.perturb_std_birth > 0:
            raise ValueError("perturb_std must be positive")
        self._repr_args = dict(name=name, **hyper, perturb_std=perturb_std, perturb_std_birth=perturb_std_birth)

    @property
    def name(self) -> str:
        return self._name

    def get_perturb_std(self, position=None):
        return self.perturb_std

    def get_perturb_std_birth(self, position=None):
        return self.perturb_std_birth

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
        self.vmin = _scalar("vmin", vmin)
        self.vmax = _scalar("vmax", vmax)
        if not self.vmin < self.vmax:
            raise ValueError("vmin should be less than vmax")
        self.delta = self.vmax - self.vmin

    def get_vmin_vmax(self, position=None):
        return self.vmin, self.vmax

    def get_delta(self, position=None):
        return self.delta

    def log_prior(self, value, position=None):
        return -math.log(self.vmax - self.vmin) if self.vmin <= value <= self.vmax else -math.inf

    def _device_params(self):
        return "uniform", self.vmin, self.vmax


class GaussianPrior(Prior):
    """``GaussianPrior(name, mean, std, perturb_std, perturb_std_birth=None, position=None)``
    (reference prior/_prior.py:484-682)."""

    kind = "gaussian"

    def __init__(self, name, mean, std, perturb_std, perturb_std_birth=None, position=None):
        super().__init__(name, perturb_std, perturb_std_birth, position, mean=mean, std=std)
        self.mean = _scalar("mean", mean)
        self.std = _scalar("std", std)
        if not self.std > 0:
            raise ValueError("std must be positive")

    def get_mean(self, position=None):
        return self.mean

    def get_std(self, position=None):
        return self.std

    def log_prior(self, value, position=None):
        return -0.5 * math.log(2 * math.pi) - math.log(self.std) - 0.5 * ((value - self.mean) / self.std) ** 2

    def _device_params(self):
        return "gaussian", self.mean, self.std


class LaplacePrior(Prior):
    """``LaplacePrior(name, mean, scale, perturb_std, perturb_std_birth=None, position=None)``
    (reference prior/_prior.py:685-881)."""

    kind = "laplace"

    def __init__(self, name, mean, scale, perturb_std, perturb_std_birth=None, position=None):
        super().__init__(name, perturb_std, perturb_std_birth, position, mean=mean, scale=scale)
        self.mean = _scalar("mean", mean)
        self.scale = _scalar("scale", scale)
        if not self.scale > 0:
            raise ValueError("scale must be positive")

    def get_mean(self, position=None):
        return self.mean

    def get_scale(self, position=None):
        return self.scale

    def log_prior(self, value, position=None):
        return -math.log(2) - math.log(self.scale) - abs(value - self.mean) / self.scale

    def _device_params(self):
        return "laplace", self.mean, self.scale


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
