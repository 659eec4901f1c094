"""
Host-side description of one (model, series, objective) problem and the
introspection that derives it from an objective object.

The reference has no evaluator plug-in hook and no operator registry
(SURVEY.md section 8b): the function handed to an evaluator is an object graph

    [ProbabilityBasedError] -> [Transformed{ErrorMeasure,LogLikelihood,LogPDF}]
        -> [LogPosterior] -> SumOfSquaresError |
        GaussianKnownSigmaLogLikelihood | GaussianLogLikelihood
        -> SingleOutputProblem | MultiOutputProblem -> toy model

``describe()`` peels that graph, for reference objects and for the mirror
classes of this package alike (both keep their state in the same private
attributes), into a ``ProblemSpec`` whose fields are exactly the arguments of
``pb200_problem_create``.  Anything it does not recognise raises ``TypeError``:
there is deliberately no "call the Python function instead" path.
"""
import numpy as np

from . import _lib

_OWNERS = ('pints', 'pints_b200')


def _owned(obj):
    mod = type(obj).__module__ or ''
    return mod.split('.')[0] in _OWNERS


def _kind(obj):
    """Class name of objects defined by the reference or by this package; a
    user subclass (which may override anything) is not trusted."""
    return type(obj).__name__ if _owned(obj) else None


class ProblemSpec(object):
    """Plain-data problem description (host arrays only)."""

    def __init__(self, model, model_consts, times, values, objective,
                 obj_consts, sign=1.0, prior_lower=None, prior_upper=None,
                 log_prior=0.0, multi_output=None, prior_terms=None):
        self.model = int(model)
        #: host-side transformation of the positions (set by describe())
        self.transform = None
        self.model_consts = [float(v) for v in model_consts]
        self.times = np.ascontiguousarray(times, dtype=np.float64)
        values = np.asarray(values, dtype=np.float64)
        if multi_output is None:
            multi_output = values.ndim == 2
        self.multi_output = bool(multi_output)
        if values.ndim == 1:
            values = values.reshape((-1, 1))
        self.values = np.ascontiguousarray(values)
        self.n_times = int(self.times.shape[0])
        self.n_outputs = int(self.values.shape[1])
        if self.values.shape[0] != self.n_times:
            raise ValueError('Times and values arrays must have same length.')
        self.objective = int(objective)
        self.obj_consts = [float(v) for v in obj_consts]
        self.sign = float(sign)
        self.prior_lower = None if prior_lower is None else \
            [float(v) for v in prior_lower]
        self.prior_upper = None if prior_upper is None else \
            [float(v) for v in prior_upper]
        self.log_prior = float(log_prior)
        # composed prior: [(kind, index, c0, c1, c2)] in the sub-priors' order
        self.prior_terms = [tuple(t) for t in prior_terms] if prior_terms \
            else None

    # ------------------------------------------------------------------
    def n_model_parameters(self):
        fixed = {_lib.MODEL_LOGISTIC: 2, _lib.MODEL_HH_IK: 5,
                 _lib.MODEL_FHN: 3, _lib.MODEL_LOTKA_VOLTERRA: 4,
                 _lib.MODEL_SIR: 3, _lib.MODEL_GOODWIN: 5,
                 _lib.MODEL_HES1: 4, _lib.MODEL_REPRESSILATOR: 4,
                 _lib.MODEL_BEELER_REUTER: 5}
        return fixed.get(self.model, self.n_outputs)

    def n_parameters(self):
        extra = self.n_outputs if self.objective == _lib.OBJ_GAUSSIAN else 0
        return self.n_model_parameters() + extra

    def is_ode(self):
        return self.model in (_lib.MODEL_FHN, _lib.MODEL_LOTKA_VOLTERRA,
                              _lib.MODEL_SIR, _lib.MODEL_GOODWIN,
                              _lib.MODEL_HES1, _lib.MODEL_REPRESSILATOR,
                              _lib.MODEL_BEELER_REUTER)

    def packed_series(self):
        """Rows ``[t_j, v_j0 ..]``, then ``SENTINEL_ROWS`` rows with time
        ``+inf``, padded to whole 16-byte units: the layout
        ``pb200_problem_create`` expects in device memory."""
        row = 1 + self.n_outputs
        nt, extra = self.n_times, _lib.SENTINEL_ROWS
        n = (nt + extra) * row
        out = np.zeros((n + 1) & ~1, dtype=np.float64)
        block = out[:n].reshape((nt + extra, row))
        block[:nt, 0] = self.times
        block[:nt, 1:] = self.values
        block[nt:, 0] = np.inf
        return out

    def flops_per_eval(self, steps):
        """Algorithmic FP64 flop of one evaluation given the attempted RK
        steps: FHN 234 S + 24 n_t, LV 222 S + 24 n_t,
        SIR 298 S + 32 n_t; closed-form models 8 n_t / 12 n_t + 750."""
        nt = self.n_times
        steps = np.asarray(steps, dtype=np.float64)
        if self.model == _lib.MODEL_FHN:
            return 234.0 * steps + 24.0 * nt
        if self.model == _lib.MODEL_LOTKA_VOLTERRA:
            return 222.0 * steps + 24.0 * nt
        if self.model == _lib.MODEL_SIR:
            return 298.0 * steps + 32.0 * nt
        # same accounting for the other ODE models: 64 per state and step for
        # the stage sums, error estimate and norm; 6 right-hand sides (pow and
        # division count 1); 10 for the controller; 18 per state for the dense
        # output; per observation 2 + 11 per output
        if self.model == _lib.MODEL_GOODWIN:
            return (3 * 64 + 6 * 16 + 10 + 54) * steps + 35.0 * nt
        if self.model == _lib.MODEL_HES1:
            return (3 * 64 + 6 * 14 + 10 + 54) * steps + 13.0 * nt
        if self.model == _lib.MODEL_REPRESSILATOR:
            return (6 * 64 + 6 * 27 + 10 + 108) * steps + 35.0 * nt
        if self.model == _lib.MODEL_BEELER_REUTER:
            # Rosenbrock 4(3) attempt: three right-hand sides (27 exp + 1 log,
            # counted 1 each as the reference writes them, + 19 divisions +
            # 106 add/mul = 153), analytic V column 90, other Jacobian pieces and
            # elimination factors 110,
            # four arrow solves 180, stage sums, new state and error 208,
            # Hermite coefficients 20; per observation 2 + 8 per output
            return (3 * 153 + 90 + 518) * steps + 18.0 * nt
        if self.model == _lib.MODEL_HH_IK:
            return 0.0 * steps + 12.0 * nt + 750.0
        return 0.0 * steps + 8.0 * nt * self.n_outputs


# ----------------------------------------------------------------------
# model introspection
# ----------------------------------------------------------------------

def describe_model(model):
    """(model id, constants) of a toy model object; ``TypeError`` otherwise."""
    kind = _kind(model)
    if kind == 'LogisticModel':
        return _lib.MODEL_LOGISTIC, [model._p0]
    if kind == 'HodgkinHuxleyIKModel':
        events = np.asarray(model._events, dtype=float)
        volts = np.asarray(model._voltages, dtype=float)
        if len(events) != len(volts) or len(events) > _lib.MAX_EVENTS:
            raise TypeError('unsupported voltage protocol')
        return _lib.MODEL_HH_IK, [model._n0, model._g_max, model._E_k,
                                  model._v_hold, len(events)] \
            + list(events) + list(volts)
    if kind == 'ConstantModel':
        n = int(model._n)
        if n > _lib.MAX_OUTPUTS:
            raise TypeError('ConstantModel with more than %d outputs'
                            % _lib.MAX_OUTPUTS)
        if not np.array_equal(np.asarray(model._r), np.arange(1, 1 + n)):
            raise TypeError('unexpected ConstantModel multipliers')
        return _lib.MODEL_CONSTANT, []
    if kind == 'FitzhughNagumoModel':
        y0 = np.asarray(model._y0, dtype=float)
        return _lib.MODEL_FHN, [y0[0], y0[1]]
    if kind == 'LotkaVolterraModel':
        y0 = np.asarray(model._y0, dtype=float)
        return _lib.MODEL_LOTKA_VOLTERRA, [y0[0], y0[1]]
    if kind == 'SIRModel':
        y0 = np.asarray(model._y0)
        # an integer y0 (the default) truncates S0 on assignment
        truncate = 1.0 if y0.dtype.kind in 'iu' else 0.0
        return _lib.MODEL_SIR, [float(y0[0]), float(y0[1]), float(y0[2]),
                                truncate]
    if kind == 'GoodwinOscillatorModel':
        y0 = np.asarray(model._y0, dtype=float)
        return _lib.MODEL_GOODWIN, [y0[0], y0[1], y0[2]]
    if kind == 'Hes1Model':
        y0 = np.asarray(model._y0, dtype=float)
        return _lib.MODEL_HES1, [y0[0], y0[1], y0[2], float(model._kdeg)]
    if kind == 'RepressilatorModel':
        y0 = np.asarray(model._y0, dtype=float)
        return _lib.MODEL_REPRESSILATOR, list(y0)
    if kind == 'ActionPotentialModel':
        return _lib.MODEL_BEELER_REUTER, [float(v) for v in (
            model._v0, model._cai0, model._m0, model._h0, model._j0,
            model._d0, model._f0, model._x10, model._C_m, model._E_Na,
            model._I_Stim_amp, model._I_Stim_period, model._I_Stim_length)]
    raise TypeError(
        'CudaEvaluator supports the toy models LogisticModel, '
        'HodgkinHuxleyIKModel, ConstantModel, FitzhughNagumoModel, '
        'LotkaVolterraModel, SIRModel, GoodwinOscillatorModel, Hes1Model, '
        'RepressilatorModel and ActionPotentialModel (Beeler-Reuter); '
        'got %s.%s (no CPU fallback).'
        % (type(model).__module__, type(model).__name__))


def describe_prior(prior):
    """(lower, upper, constant, terms) of a supported prior: a UniformLogPrior
    over RectangularBoundaries, a GaussianLogPrior, or a ComposedLogPrior of
    those (pints/_log_priors.py:170-219, :441-468, :1264-1320)."""
    kind = _kind(prior)
    if kind == 'UniformLogPrior':
        if _kind(prior._boundaries) != 'RectangularBoundaries':
            raise TypeError('only UniformLogPrior over RectangularBoundaries '
                            'is fused into the CUDA path')
        return (np.asarray(prior._boundaries.lower(), dtype=float),
                np.asarray(prior._boundaries.upper(), dtype=float),
                float(prior._value), None)
    if kind == 'GaussianLogPrior':
        subs = [prior]
    elif kind == 'ComposedLogPrior':
        subs = list(prior._priors)
    else:
        raise TypeError(
            'only UniformLogPrior, GaussianLogPrior and ComposedLogPrior of '
            'those are fused into the CUDA path; got %s' % type(prior).__name__)
    lower, upper, terms = [], [], []
    for sub in subs:
        k = _kind(sub)
        if k == 'UniformLogPrior' and \
                _kind(sub._boundaries) == 'RectangularBoundaries':
            lower += list(np.asarray(sub._boundaries.lower(), dtype=float))
            upper += list(np.asarray(sub._boundaries.upper(), dtype=float))
            terms.append((0, 0, float(sub._value), 0.0, 0.0))
        elif k == 'GaussianLogPrior':
            terms.append((1, len(lower), float(sub._offset),
                          float(sub._factor), float(sub._mean)))
            lower.append(-np.inf)
            upper.append(np.inf)
        else:
            raise TypeError('unsupported sub-prior %s' % type(sub).__name__)
    if len(terms) > _lib.MAX_PARAMS:
        raise TypeError('too many sub-priors')
    return np.array(lower), np.array(upper), 0.0, terms


def describe(function):
    """``ProblemSpec`` of an objective object graph; ``TypeError`` if any part
    of it is outside the supported hot path."""
    sign = 1.0
    prior = None
    transform = None           # (transformation, add the Jacobian term?, sign outside)
    f = function
    for _ in range(8):
        kind = _kind(f)
        if kind == 'ProbabilityBasedError':
            sign = -sign
            f = f._log_pdf
        elif kind in ('TransformedErrorMeasure', 'TransformedLogLikelihood',
                      'TransformedLogPDF'):
            # f(to_model(q)) [+ log |det J(q)|]: the rows are mapped on the
            # host before the launch (pints_b200/_transform.py)
            if transform is not None or prior is not None:
                raise TypeError('only one transformation, outside any '
                                'LogPosterior, is supported')
            transform = (f._transform, kind == 'TransformedLogPDF', sign)
            f = getattr(f, {'TransformedErrorMeasure': '_error',
                            'TransformedLogLikelihood': '_log_likelihood',
                            'TransformedLogPDF': '_log_pdf'}[kind])
        elif kind == 'LogPosterior':
            if prior is not None:
                raise TypeError('nested LogPosterior is not supported')
            prior = f._log_prior
            f = f._log_likelihood
        else:
            break
    kind = _kind(f)
    if kind not in ('SumOfSquaresError', 'MeanSquaredError',
                    'RootMeanSquaredError', 'NormalisedRootMeanSquaredError',
                    'GaussianKnownSigmaLogLikelihood',
                    'GaussianLogLikelihood'):
        raise TypeError(
            'CudaEvaluator supports SumOfSquaresError, MeanSquaredError, '
            '(Normalised)RootMeanSquaredError, '
            'GaussianKnownSigmaLogLikelihood and GaussianLogLikelihood '
            '(optionally inside LogPosterior with a UniformLogPrior and/or '
            'ProbabilityBasedError); got %s.%s (no CPU fallback).'
            % (type(f).__module__, type(f).__name__))
    problem = f._problem
    if _kind(problem) not in ('SingleOutputProblem', 'MultiOutputProblem'):
        raise TypeError('unsupported problem class %s'
                        % type(problem).__name__)
    multi = _kind(problem) == 'MultiOutputProblem'
    model_id, model_consts = describe_model(problem._model)
    times = np.asarray(problem._times, dtype=float)
    values = np.asarray(problem._values, dtype=float)
    n_outputs = values.shape[1] if values.ndim == 2 else 1
    n_times = len(times)
    if kind == 'SumOfSquaresError':
        objective = _lib.OBJ_SSE
        obj_consts = list(np.asarray(f._weights, dtype=float))
    elif kind == 'MeanSquaredError':
        objective = _lib.OBJ_MSE
        obj_consts = list(np.asarray(f._weights, dtype=float)) + \
            [float(f._ninv)]
    elif kind in ('RootMeanSquaredError', 'NormalisedRootMeanSquaredError'):
        if multi:
            raise TypeError('%s needs a SingleOutputProblem' % kind)
        objective = _lib.OBJ_RMSE
        obj_consts = [float(f._ninv), float(getattr(f, '_norm', 1.0))]
    elif kind == 'GaussianKnownSigmaLogLikelihood':
        objective = _lib.OBJ_GAUSSIAN_KNOWN_SIGMA
        offset = np.ones(n_outputs) * np.asarray(f._offset, dtype=float)
        multip = np.ones(n_outputs) * np.asarray(f._multip, dtype=float)
        obj_consts = list(offset) + list(multip)
    else:
        objective = _lib.OBJ_GAUSSIAN
        obj_consts = [float(f._logn)]
        if int(f._nt) != n_times or int(f._no) != n_outputs:
            raise TypeError('inconsistent GaussianLogLikelihood')
    lower = upper = None
    log_prior = 0.0
    terms = None
    if prior is not None:
        lower, upper, log_prior, terms = describe_prior(prior)
    spec = ProblemSpec(model_id, model_consts, times, values, objective,
                       obj_consts, sign, lower, upper, log_prior,
                       multi_output=multi, prior_terms=terms)
    if lower is not None and len(lower) != spec.n_parameters():
        raise TypeError('prior and likelihood dimensions differ')
    if transform is not None and \
            transform[0].n_parameters() != spec.n_parameters():
        raise TypeError('transformation and objective dimensions differ')
    spec.transform = transform
    return spec
