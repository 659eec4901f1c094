"""
Mirror of the reference's problem layer (pints/_core.py:12-327) with a batched
``evaluate``: ``problem.evaluate(x)`` for one vector returns what the reference
returns, ``problem.evaluate(xs)`` for ``(N, d)`` returns ``(N, n_times[,
n_outputs])`` -- both computed on the GPU through ``pb200_simulate``.
"""
import numpy as np

from ._compat import base
from ._util import is_batch, matrix2d, vector


class ForwardModel(base('ForwardModel')):
    """Mirror of ``pints.ForwardModel`` (pints/_core.py:12-56)."""

    def __init__(self):
        pass

    def n_parameters(self):
        raise NotImplementedError

    def simulate(self, parameters, times):
        raise NotImplementedError

    def n_outputs(self):
        return 1


class SingleOutputProblem(base('SingleOutputProblem')):
    """Mirror of ``pints.SingleOutputProblem`` (pints/_core.py:102-210):
    times non-negative and strictly increasing, one value per time."""

    def __init__(self, model, times, values):
        self._model = model
        if model.n_outputs() != 1:
            raise ValueError(
                'Only single-output models can be used for a'
                ' SingleOutputProblem.')
        self._times = vector(times)
        if np.any(self._times < 0):
            raise ValueError('Times can not be negative.')
        if np.any(self._times[:-1] >= self._times[1:]):
            raise ValueError('Times must be increasing.')
        self._values = vector(values)
        self._n_parameters = int(model.n_parameters())
        self._n_times = len(self._times)
        if len(self._values) != self._n_times:
            raise ValueError(
                'Times and values arrays must have same length.')

    def evaluate(self, parameters):
        y = np.asarray(self._model.simulate(parameters, self._times))
        if is_batch(parameters):
            return y.reshape((len(parameters), self._n_times))
        return y.reshape((self._n_times,))

    def model(self):
        return self._model

    def n_outputs(self):
        return 1

    def n_parameters(self):
        return self._n_parameters

    def n_times(self):
        return self._n_times

    def times(self):
        return self._times

    def values(self):
        return self._values


class MultiOutputProblem(base('MultiOutputProblem')):
    """Mirror of ``pints.MultiOutputProblem`` (pints/_core.py:213-327): times
    non-negative and non-decreasing, values ``(n_times, n_outputs)``."""

    def __init__(self, model, times, values):
        self._model = model
        self._times = vector(times)
        if np.any(self._times < 0):
            raise ValueError('Times cannot be negative.')
        if np.any(self._times[:-1] > self._times[1:]):
            raise ValueError('Times must be non-decreasing.')
        self._values = matrix2d(values)
        self._n_parameters = int(model.n_parameters())
        self._n_outputs = int(model.n_outputs())
        self._n_times = len(self._times)
        if self._values.shape != (self._n_times, self._n_outputs):
            raise ValueError(
                'Values array must have shape `(n_times, n_outputs)`.')

    def evaluate(self, parameters):
        y = np.asarray(self._model.simulate(parameters, self._times))
        if is_batch(parameters):
            return y.reshape(len(parameters), self._n_times, self._n_outputs)
        return y.reshape(self._n_times, self._n_outputs)

    def model(self):
        return self._model

    def n_outputs(self):
        return self._n_outputs

    def n_parameters(self):
        return self._n_parameters

    def n_times(self):
        return self._n_times

    def times(self):
        return self._times

    def values(self):
        return self._values
