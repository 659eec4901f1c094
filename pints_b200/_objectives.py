"""
Mirror of the objective layer on the hot path with batched ``__call__``:

  SumOfSquaresError                 pints/_error_measures.py:338-385
  ProbabilityBasedError             pints/_error_measures.py:164-198
  GaussianKnownSigmaLogLikelihood   pints/_log_likelihoods.py:989-1062
  GaussianLogLikelihood             pints/_log_likelihoods.py:1065-1158
  LogPosterior                      pints/_log_pdfs.py:164-241
  UniformLogPrior                   pints/_log_priors.py:1264-1320
  RectangularBoundaries             pints/_boundaries.py:100-239

``f(x)`` with one parameter vector returns a float, ``f(xs)`` with ``(N, d)``
returns an ``(N,)`` array; both run on the GPU through ``CudaEvaluator``
(the constants are pre-computed on the host exactly as the reference does, so
the same numbers reach the kernel).
"""
import numpy as np

from ._compat import base, is_instance
from ._util import is_batch, vector


class _DeviceCallable(object):
    """Mixin: ``__call__`` through a cached CudaEvaluator."""

    _evaluator = None

    def _device_evaluator(self):
        if self._evaluator is None:
            from ._evaluation import CudaEvaluator
            self._evaluator = CudaEvaluator(self, as_list=False)
        return self._evaluator

    def __call__(self, x):
        ev = self._device_evaluator()
        if is_batch(x):
            return ev.evaluate_array(x)
        x = np.asarray(x, dtype=float).reshape((1, -1))
        return float(ev.evaluate_array(x)[0])


# ---------------------------------------------------------------- errors

class ErrorMeasure(base('ErrorMeasure')):
    """Mirror of ``pints.ErrorMeasure`` (pints/_error_measures.py:11-47)."""

    def __call__(self, x):
        raise NotImplementedError

    def n_parameters(self):
        raise NotImplementedError


class ProblemErrorMeasure(ErrorMeasure):
    """pints/_error_measures.py:50-71"""

    def __init__(self, problem):
        self._problem = problem
        self._times = problem.times()
        self._values = problem.values()
        self._n_outputs = problem.n_outputs()
        self._n_parameters = problem.n_parameters()
        self._n_times = len(self._times)

    def n_parameters(self):
        return self._n_parameters

    def problem(self):
        return self._problem

    def _problem_is_multi_output(self):
        return np.ndim(self._values) == 2


class SumOfSquaresError(_DeviceCallable, ProblemErrorMeasure):
    """``sum_o w_o sum_t (sim - data)^2`` (pints/_error_measures.py:338-385)."""

    def __init__(self, problem, weights=None):
        super().__init__(problem)
        if weights is None:
            weights = [1] * self._n_outputs
        elif self._n_outputs != len(weights):
            raise ValueError(
                'Number of weights must match number of problem outputs.')
        self._weights = np.asarray([float(w) for w in weights])


class MeanSquaredError(_DeviceCallable, ProblemErrorMeasure):
    """``sum_o w_o sum_t (sim - data)^2 / (n_t n_o)``
    (pints/_error_measures.py:74-114)."""

    def __init__(self, problem, weights=None):
        super().__init__(problem)
        self._ninv = 1.0 / np.prod(self._values.shape)
        if weights is None:
            weights = [1] * self._n_outputs
        elif self._n_outputs != len(weights):
            raise ValueError(
                'Number of weights must match number of problem outputs.')
        self._weights = np.asarray([float(w) for w in weights])


class RootMeanSquaredError(_DeviceCallable, ProblemErrorMeasure):
    """``sqrt(sum_t (sim - data)^2 / n_t)``, single-output problems only
    (pints/_error_measures.py:201-229)."""

    def __init__(self, problem):
        super().__init__(problem)
        if self._problem_is_multi_output():
            raise ValueError(
                'This measure is only defined for single output problems.')
        self._ninv = 1.0 / len(self._values)


class NormalisedRootMeanSquaredError(_DeviceCallable, ProblemErrorMeasure):
    """``sqrt(mean((sim - data)^2)) / sqrt(mean(data^2))``
    (pints/_error_measures.py:135-161)."""

    def __init__(self, problem):
        super().__init__(problem)
        if self._problem_is_multi_output():
            raise ValueError(
                'This measure is only defined for single output problems.')
        self._ninv = 1.0 / len(self._values)
        self._norm = 1.0 / np.sqrt(self._ninv * np.sum(self._values**2))


class ProbabilityBasedError(_DeviceCallable, ErrorMeasure):
    """``-log_pdf(x)`` (pints/_error_measures.py:164-198)."""

    def __init__(self, log_pdf):
        if not is_instance(log_pdf, LogPDF, 'LogPDF'):
            raise ValueError(
                'Given log_pdf must be an instance of pints.LogPDF.')
        self._log_pdf = log_pdf

    def n_parameters(self):
        return self._log_pdf.n_parameters()


# ------------------------------------------------------------- log pdfs

class LogPDF(base('LogPDF')):
    """Mirror of ``pints.LogPDF`` (pints/_log_pdfs.py:11-69)."""

    def __call__(self, x):
        raise NotImplementedError

    def n_parameters(self):
        raise NotImplementedError


class LogPrior(LogPDF, base('LogPrior')):
    """pints/_log_pdfs.py:72-118"""


class LogLikelihood(LogPDF, base('LogLikelihood')):
    """pints/_log_pdfs.py:121-131"""


class ProblemLogLikelihood(LogLikelihood, base('ProblemLogLikelihood')):
    """pints/_log_pdfs.py:134-161"""

    def __init__(self, problem):
        self._problem = problem
        self._values = problem.values()
        self._times = problem.times()
        self._n_parameters = problem.n_parameters()

    def n_parameters(self):
        return self._n_parameters

    def problem(self):
        return self._problem


class GaussianKnownSigmaLogLikelihood(_DeviceCallable, ProblemLogLikelihood):
    """Independent Gaussian noise with known standard deviation(s)
    (pints/_log_likelihoods.py:989-1062)."""

    def __init__(self, problem, sigma):
        super().__init__(problem)
        self._no = problem.n_outputs()
        self._np = problem.n_parameters()
        self._nt = problem.n_times()
        if np.isscalar(sigma):
            sigma = np.ones(self._no) * float(sigma)
        else:
            sigma = vector(sigma)
            if len(sigma) != self._no:
                raise ValueError(
                    'Sigma must be a scalar or a vector of length n_outputs.')
        if np.any(sigma <= 0):
            raise ValueError('Standard deviation must be greater than zero.')
        # constants, evaluated in the reference's order (:1027-1032)
        self._offset = -0.5 * self._nt * np.log(2 * np.pi)
        self._offset = self._offset - self._nt * np.log(sigma)
        self._multip = -1 / (2.0 * sigma**2)


class GaussianLogLikelihood(_DeviceCallable, ProblemLogLikelihood):
    """Independent Gaussian noise whose standard deviation(s) are the last
    ``n_outputs`` parameters; ``sigma <= 0`` scores ``-inf``
    (pints/_log_likelihoods.py:1065-1158)."""

    def __init__(self, problem):
        super().__init__(problem)
        self._nt = len(self._times)
        self._no = problem.n_outputs()
        self._n_parameters = problem.n_parameters() + self._no
        self._logn = 0.5 * self._nt * np.log(2 * np.pi)


# ----------------------------------------------------- prior / posterior

class RectangularBoundaries(base('RectangularBoundaries')):
    """``lower <= x < upper`` (pints/_boundaries.py:100-239)."""

    def __init__(self, lower, upper):
        self._lower = vector(lower)
        self._upper = vector(upper)
        self._n_parameters = len(self._lower)
        if len(self._upper) != self._n_parameters:
            raise ValueError('Lower and upper bounds must have same length.')
        if self._n_parameters < 1:
            raise ValueError('The parameter space must have a dimension > 0')
        if not np.all(self._upper > self._lower):
            raise ValueError('Upper bounds must exceed lower bounds.')

    def check(self, parameters):
        """One vector -> bool; ``(N, d)`` -> bool array."""
        p = np.asarray(parameters)
        axis = -1
        out = ~(np.any(p < self._lower, axis=axis)
                | np.any(p >= self._upper, axis=axis))
        return bool(out) if p.ndim == 1 else out

    def n_parameters(self):
        return self._n_parameters

    def lower(self):
        return self._lower

    def upper(self):
        return self._upper

    def range(self):
        return self._upper - self._lower

    def sample(self, n=1):
        return np.random.uniform(self._lower, self._upper,
                                 size=(n, self._n_parameters))


class UniformLogPrior(LogPrior):
    """``-log(volume)`` inside the box, ``-inf`` outside
    (pints/_log_priors.py:1264-1320).  Evaluated on the host when called on
    its own; fused into the kernel epilogue inside a LogPosterior."""

    def __init__(self, lower_or_boundaries, upper=None):
        if upper is None:
            b = lower_or_boundaries
            if type(b).__name__ != 'RectangularBoundaries':
                raise ValueError(
                    'UniformLogPrior requires a lower and an upper bound, or a'
                    ' single RectangularBoundaries object.')
            self._boundaries = b
        else:
            self._boundaries = RectangularBoundaries(lower_or_boundaries,
                                                     upper)
        self._n_parameters = self._boundaries.n_parameters()
        self._value = -np.log(np.prod(self._boundaries.range()))

    def __call__(self, x):
        inside = self._boundaries.check(x)
        if np.ndim(inside) == 0:
            return self._value if inside else -np.inf
        return np.where(inside, self._value, -np.inf)

    def n_parameters(self):
        return self._n_parameters

    def sample(self, n=1):
        return self._boundaries.sample(n)


class GaussianLogPrior(LogPrior):
    """One-dimensional normal prior (pints/_log_priors.py:441-483)."""

    def __init__(self, mean, sd):
        self._mean = float(mean)
        self._sd = float(sd)
        if self._sd <= 0:
            raise ValueError('sd parameter must be positive')
        self._offset = np.log(1 / np.sqrt(2 * np.pi * self._sd ** 2))
        self._factor = 1 / (2 * self._sd ** 2)
        self._factor2 = 1 / self._sd**2

    def __call__(self, x):
        return self._offset - self._factor * (x[0] - self._mean)**2

    def n_parameters(self):
        return 1

    def mean(self):
        return self._mean

    def sample(self, n=1):
        return np.random.normal(self._mean, self._sd, size=(n, 1))


class ComposedLogPrior(LogPrior):
    """Independent sub-priors side by side (pints/_log_priors.py:170-275)."""

    def __init__(self, *priors):
        if len(priors) < 1:
            raise ValueError('Must have at least one sub-prior')
        self._n_parameters = 0
        for prior in priors:
            if not is_instance(prior, LogPrior, 'LogPrior'):
                raise ValueError('All sub-priors must extend pints.LogPrior.')
            self._n_parameters += prior.n_parameters()
        self._priors = priors

    def __call__(self, x):
        output = 0
        lo = hi = 0
        for prior in self._priors:
            lo = hi
            hi += prior.n_parameters()
            output += prior(x[lo:hi])
        return output

    def n_parameters(self):
        return self._n_parameters

    def sample(self, n=1):
        output = np.zeros((n, self._n_parameters))
        lo = hi = 0
        for prior in self._priors:
            lo = hi
            hi += prior.n_parameters()
            output[:, lo:hi] = prior.sample(n)
        return output


class LogPosterior(_DeviceCallable, LogPDF):
    """``log_prior(x) + log_likelihood(x)``; a point outside the prior's
    support is not simulated (pints/_log_pdfs.py:164-241)."""

    def __init__(self, log_likelihood, log_prior):
        if not is_instance(log_prior, LogPrior, 'LogPrior'):
            raise ValueError('Given prior must extend pints.LogPrior.')
        if not isinstance(log_likelihood,
                          (LogLikelihood, base('LogLikelihood'))):
            raise ValueError(
                'Given log_likelihood must extend pints.LogLikelihood.')
        self._n_parameters = log_prior.n_parameters()
        if log_likelihood.n_parameters() != self._n_parameters:
            raise ValueError(
                'Given log_prior and log_likelihood must have same dimension.')
        self._log_prior = log_prior
        self._log_likelihood = log_likelihood
        self._minf = -np.inf

    def log_likelihood(self):
        return self._log_likelihood

    def log_prior(self):
        return self._log_prior

    def n_parameters(self):
        return self._n_parameters
