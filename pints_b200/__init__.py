"""
pints_b200 -- a B200-native batched evaluation engine for PINTS.

One hot path, built from scratch for sm_100a: scoring a whole optimiser
population or a whole set of MCMC chains as a single fused forward solve plus
likelihood reduction.  The public surface mirrors the
reference's own interfaces for that path:

  CudaEvaluator                       drop-in for Sequential/ParallelEvaluator
  SingleOutputProblem, MultiOutputProblem
  toy.LogisticModel, toy.FitzhughNagumoModel, toy.HodgkinHuxleyIKModel,
  toy.LotkaVolterraModel, toy.SIRModel, toy.ConstantModel
  SumOfSquaresError, GaussianKnownSigmaLogLikelihood, GaussianLogLikelihood,
  ProbabilityBasedError, LogPosterior, UniformLogPrior, RectangularBoundaries
  HaarioBardenetACMC, DifferentialEvolutionMCMC   device-resident samplers
  XNES (vectorised), DeviceXNES, DeviceSNES, CudaOptimisationController,
  CudaMCMCController,
  cuda_evaluators()                   run the reference's own controllers on GPU

Host code is Python; compute is hand-written CUDA behind the C ABI of
``include/pints_b200.h`` (bound with ctypes in ``_lib``); PyTorch supplies
device buffers, streams and ``torch.distributed`` only.  There is no CPU
fallback: without the built library and a CUDA device, evaluation raises.
"""
from ._util import vector, matrix2d                                  # noqa
from ._spec import ProblemSpec, describe, describe_model             # noqa
from ._core import (ForwardModel, SingleOutputProblem,               # noqa
                    MultiOutputProblem)
from ._objectives import (ErrorMeasure, ProblemErrorMeasure,         # noqa
                          SumOfSquaresError, MeanSquaredError,
                          RootMeanSquaredError,
                          NormalisedRootMeanSquaredError,
                          ProbabilityBasedError, LogPDF,
                          LogPrior, LogLikelihood, ProblemLogLikelihood,
                          GaussianKnownSigmaLogLikelihood,
                          GaussianLogLikelihood, RectangularBoundaries,
                          UniformLogPrior, GaussianLogPrior,
                          ComposedLogPrior, LogPosterior)
from . import toy                                                    # noqa

__version__ = '0.1.0'

_LAZY = {
    'Evaluator': '_evaluation', 'CudaEvaluator': '_evaluation',
    'pinned_empty': '_evaluation', 'ScoreList': '_evaluation',
    'DeviceProblem': '_engine', 'fp64_probe': '_engine',
    'require_cuda': '_engine',
    'ShardedEvaluator': '_dist', 'ScoreExchange': '_dist',
    'HaarioBardenetACMC': '_samplers', 'HaarioACMC': '_samplers',
    'RaoBlackwellACMC': '_samplers',
    'MetropolisRandomWalkMCMC': '_samplers',
    'DifferentialEvolutionMCMC': '_samplers',
    'XNES': '_controllers', 'SNES': '_controllers', 'PSO': '_controllers',
    'BareCMAES': '_controllers', 'cuda_evaluators': '_controllers',
    'DeviceXNES': '_device_optimisers', 'DeviceSNES': '_device_optimisers',
    'DeviceArgsort': '_device_optimisers',
    'CudaOptimisationController': '_controllers',
    'CudaMCMCController': '_controllers',
}


def __getattr__(name):
    # torch-dependent parts load on first use so that `import pints_b200`
    # stays light
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError('module %r has no attribute %r'
                             % (__name__, name))
    import importlib
    value = getattr(importlib.import_module('.' + mod, __name__), name)
    globals()[name] = value
    return value
