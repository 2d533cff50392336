"""bayesbay_b200 -- B200-native drop-in for BayesBay's batched reversible-jump MCMC hot path.

Same public names as ``bayesbay`` 0.3.11 (src/bayesbay/__init__.py:3-55); the chain step runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``libbbb200.so`` (``include/bbb200.h``).
There is no CPU fallback: building an inversion without the compiled library or without a GPU raises.
"""
from ._version import __version__
from ._state import DataNoiseState, ParameterSpaceState, State
from . import exceptions, prior, perturbations, parameterization, discretization, likelihood, forward, samplers, synthetic
from ._markov_chain import BaseMarkovChain, MarkovChain
from ._bayes_inversion import BaseBayesianInversion, BayesianInversion
from ._batch import seed
from .likelihood import LogLikelihood, Target

__all__ = [
    "DataNoiseState", "ParameterSpaceState", "State", "Target", "LogLikelihood", "MarkovChain", "BaseMarkovChain",
    "BayesianInversion", "BaseBayesianInversion", "seed",
]
