"""``Target`` and ``LogLikelihood`` descriptors (reference src/bayesbay/likelihood/)."""
from __future__ import annotations

from numbers import Number
from typing import List, Union

import numpy as np

from .forward import ForwardOperator
from .perturbations import NoisePerturbation

__all__ = ["Target", "LogLikelihood"]


class Target:
    """Observed data with fixed or hierarchical (unknown) noise level.

    ``Target(name, dobs, covariance_mat_inv=None, noise_is_correlated=False, std_min=0.01, std_max=1,
    std_perturb_std=0.1, correlation_min=0.01, correlation_max=1, correlation_perturb_std=0.1)`` --
    reference likelihood/_target.py:10-208.  Full (2-D) inverse covariance matrices are not supported."""

    def __init__(self, name: str, dobs: np.ndarray, covariance_mat_inv: Union[Number, np.ndarray] = None,
                 noise_is_correlated: bool = False, std_min: Number = 0.01, std_max: Number = 1,
                 std_perturb_std: Number = 0.1, correlation_min: Number = 0.01, correlation_max: Number = 1,
                 correlation_perturb_std: Number = 0.1):
        self._name = name
        self.dobs = np.array(dobs)
        self.noise_is_correlated = noise_is_correlated
        self.std_min, self.std_max, self.std_perturb_std = std_min, std_max, std_perturb_std
        self.correlation_min, self.correlation_max = correlation_min, correlation_max
        self.correlation_perturb_std = correlation_perturb_std
        assert std_min <= std_max, "standard deviation minimum should be less than maximum"
        assert std_min >= 0, "standard deviation should always be positive"
        assert correlation_min <= correlation_max, "correlation minimum should be less than maximum"
        assert correlation_min >= 0, "correlation should always be positive"
        assert correlation_max <= 1, "correlation should always be less than 1"
        if noise_is_correlated and covariance_mat_inv is not None:
            raise ValueError("`noise_is_correlated` describes hierarchical noise; it cannot be combined with a fixed "
                             "`covariance_mat_inv`")
        if covariance_mat_inv is None:
            self.covariance_mat_inv = None
        elif np.isscalar(covariance_mat_inv):
            self.covariance_mat_inv = covariance_mat_inv
        else:
            self.covariance_mat_inv = np.array(covariance_mat_inv)
            if self.covariance_mat_inv.ndim not in (1, 2):
                raise ValueError("covariance_mat_inv must be a scalar, a 1-D (diagonal) or a 2-D (full) array")

    @property
    def name(self) -> str:
        return self._name

    @property
    def is_hierarchical(self) -> bool:
        """whether the data noise is unknown (i.e. to be inverted for)"""
        return self.covariance_mat_inv is None

    def __repr__(self):
        return (f"Target(name='{self.name}', dobs=ndarray of shape {self.dobs.shape}, "
                f"is_hierarchical={self.is_hierarchical})")


class LogLikelihood:
    """``LogLikelihood(targets, fwd_functions)`` (reference likelihood/_log_likelihood.py:16-332).

    ``fwd_functions`` must be device forward operators from :mod:`bayesbay_b200.forward`; the
    misfit, log-determinant and (tempered) likelihood ratio are computed by the CUDA kernels.
    ``log_like_ratio_func`` / ``log_like_func`` (arbitrary Python) are rejected: no CPU fallback."""

    def __init__(self, targets: Union[Target, List[Target]] = None, fwd_functions=None,
                 log_like_ratio_func=None, log_like_func=None):
        if log_like_ratio_func is not None or log_like_func is not None:
            raise TypeError(
                "user log-likelihood callables cannot run inside a CUDA kernel and bayesbay_b200 has no CPU "
                "fallback; provide `targets` and device forward operators (bayesbay_b200.forward) instead")
        if targets is None or fwd_functions is None:
            raise ValueError("please provide `targets` and `fwd_functions`")
        self._targets: List[Target] = []
        self._fwd_functions: list = []
        self._perturbation_funcs_observers = []
        self.log_like_ratio_func = None
        self.log_like_func = None
        self.add_targets(targets, fwd_functions)

    targets = property(lambda self: self._targets)
    fwd_functions = property(lambda self: self._fwd_functions)
    perturbation_funcs = property(lambda self: self._perturbation_funcs)
    perturbation_weights = property(lambda self: self._perturbation_weights)

    def add_targets(self, targets, fwd_functions):
        targets = targets if isinstance(targets, list) else [targets]
        fwd_functions = fwd_functions if isinstance(fwd_functions, list) else [fwd_functions]
        if len(targets) != len(fwd_functions):
            raise ValueError("the number of targets and forward functions must be equal")
        for t in targets:
            if not isinstance(t, Target):
                raise TypeError("`targets` should be a Target or a list of Targets")
        for f in fwd_functions:
            if not isinstance(f, ForwardOperator):
                raise TypeError(
                    f"`fwd_functions` got {f!r}: the device engine runs the forward operators of "
                    "bayesbay_b200.forward (RayTomographyForward, NearestLookup1D, LinearForward); arbitrary "
                    "Python callables cannot run in a CUDA kernel and there is no CPU fallback")
        self._targets.extend(targets)
        self._fwd_functions.extend(fwd_functions)
        names = [t.name for t in self._targets]
        if len(names) != len(set(names)):
            raise ValueError("duplicate target names found")
        self._init_perturbation_funcs()
        for inversion in self._perturbation_funcs_observers:
            inversion.update_log_likelihood_targets(targets)

    def add_targets_observer(self, inversion):
        self._perturbation_funcs_observers.append(inversion)

    def _init_perturbation_funcs(self):
        # one NoisePerturbation of weight 1 over all hierarchical targets (reference :293-306)
        hier = [t for t in self._targets if t.is_hierarchical]
        self._perturbation_funcs = [NoisePerturbation(hier)] if hier else []
        self._perturbation_weights = [1] if hier else []

    def __repr__(self):
        return (f"LogLikelihood(targets=[{', '.join(repr(t) for t in self.targets)}], "
                f"fwd_functions=[{', '.join(repr(f.__name__) for f in self.fwd_functions)}])")
