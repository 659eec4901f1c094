"""
The problem layer with a batched ``evaluate``: ``problem.evaluate(x)`` for one
vector returns what the reference returns (pints/_core.py:147-153, :257-265),
``problem.evaluate(xs)`` for ``(N, d)`` returns ``(N, n_times[, n_outputs])``
-- both computed on the GPU through ``pb200_simulate`` by the device models of
``pints_b200.toy``.

Where the reference is importable the classes below ARE the reference's
(``pints.SingleOutputProblem`` / ``pints.MultiOutputProblem``: constructors,
validation and accessors are inherited) with ``evaluate`` replaced by the
batch-aware one.  Without it (``pints_b200`` is standalone) a small generic
stand-in keeps the same state under the same private names -- ``_model``,
``_times``, ``_values``, ``_n_parameters``, ``_n_times``, ``_n_outputs``, which
is what ``_spec.describe`` reads from reference and stand-in objects alike --
and raises the reference's messages.
"""
import numpy as np

from ._compat import base, pints_or_none
from ._util import is_batch, matrix2d, vector


class ForwardModel(base('ForwardModel')):
    """``pints.ForwardModel`` (pints/_core.py:12-56) when the reference is
    there, else its interface: ``n_parameters()``, ``simulate(parameters,
    times)``, ``n_outputs()`` (1 unless overridden)."""

    def __init__(self):
        pass

    def n_parameters(self):
        raise NotImplementedError

    def simulate(self, parameters, times):
        raise NotImplementedError

    def n_outputs(self):
        return 1


class _BatchedEvaluate(object):
    """``evaluate`` for one parameter vector or a batch of them."""

    def evaluate(self, parameters):
        y = np.asarray(self._model.simulate(parameters, self._times))
        shape = (self._n_times,)
        if self._multi_output:
            shape += (self._n_outputs,)
        if is_batch(parameters):
            shape = (len(parameters),) + shape
        return y.reshape(shape)


_pints = pints_or_none()

if _pints is not None:

    class SingleOutputProblem(_BatchedEvaluate, _pints.SingleOutputProblem):
        """``pints.SingleOutputProblem`` with the batched ``evaluate``."""
        _multi_output = False

    class MultiOutputProblem(_BatchedEvaluate, _pints.MultiOutputProblem):
        """``pints.MultiOutputProblem`` with the batched ``evaluate``."""
        _multi_output = True

else:

    class _StandInProblem(_BatchedEvaluate):
        """No reference installed: the state ``_spec.describe`` needs, checked
        as the reference checks it (pints/_core.py:129-145, :238-255)."""
        _multi_output = False

        def __init__(self, model, times, values):
            single = not self._multi_output
            if single and model.n_outputs() != 1:
                raise ValueError('Only single-output models can be used for a'
                                 ' SingleOutputProblem.')
            t = vector(times)
            if np.any(t < 0):
                raise ValueError('Times can not be negative.' if single
                                 else 'Times cannot be negative.')
            steps = np.diff(t)
            if single and np.any(steps <= 0):
                raise ValueError('Times must be increasing.')
            if not single and np.any(steps < 0):
                raise ValueError('Times must be non-decreasing.')
            v = vector(values) if single else matrix2d(values)
            n_out = 1 if single else int(model.n_outputs())
            if single and len(v) != len(t):
                raise ValueError(
                    'Times and values arrays must have same length.')
            if not single and v.shape != (len(t), n_out):
                raise ValueError(
                    'Values array must have shape `(n_times, n_outputs)`.')
            self._model, self._times, self._values = model, t, v
            self._n_parameters = int(model.n_parameters())
            self._n_times, self._n_outputs = len(t), n_out

    # accessors: model(), times(), values(), n_parameters(), n_times(), n_outputs()
    for _name in ('model', 'times', 'values', 'n_parameters', 'n_times',
                  'n_outputs'):
        setattr(_StandInProblem, _name,
                (lambda attr: lambda self: getattr(self, attr))('_' + _name))

    class SingleOutputProblem(_StandInProblem):
        """Stand-in for ``pints.SingleOutputProblem`` (pints/_core.py:102-210)."""
        _multi_output = False

    class MultiOutputProblem(_StandInProblem):
        """Stand-in for ``pints.MultiOutputProblem`` (pints/_core.py:213-327)."""
        _multi_output = True
