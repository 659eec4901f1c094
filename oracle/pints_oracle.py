"""
CPU oracle for the pints-b200 hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain numpy/scipy restatement of what the reference
(pints-team/pints v0.6.2, mounted read-only at /root/reference while this was
written) computes on the one hot path this repository accelerates:

    Evaluator.evaluate(xs) -> f(x) -> Problem.evaluate(x) -> model.simulate(x, t)
        -> residual -> sum of squares -> Gaussian log-likelihood
    plus the ask/tell state machines of HaarioBardenetACMC and
    DifferentialEvolutionMCMC.

It is NOT part of the product: nothing under ``pints_b200/`` imports it.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or as the timed CPU baseline.

Third-party arithmetic: the reference's ODE solves happen inside
``scipy.integrate.odeint`` (LSODA), a dependency that is not vendored in the
reference and not pinned by it (``setup.py:65`` ``scipy>=0.14``).  The oracle
calls the very same routine from the very same image (scipy 1.18.1 here and on
the GPU box) with the reference's arguments (all tolerances left at their
defaults => rtol = atol = 1.49012e-8, mxstep = 500), and restates the
right-hand sides with the reference's operation order, so ``*_simulate`` below
is bit-identical to the reference on this image.

Pinning ("is the oracle the reference?"): ``tests/test_oracle_golden.py``
checks every function here against
  * the known answers held by the reference's own unit tests
    (SURVEY.md section 8c lists them with file:line), and
  * fixtures under ``tests/golden/*.npz`` that were produced by importing the
    real reference in the build container (``tests/golden/make_golden.py``).
All of those comparisons are exact (``==``) except where the reference's own
test states a number of decimal places.

Citations are relative to /root/reference.
"""
import math

import numpy as np
from scipy.integrate import odeint

TWO_PI = 2 * np.pi


# --------------------------------------------------------------------------
# helpers shared with the reference's conventions
# --------------------------------------------------------------------------

def as_vector(x):
    """1-d float64 read-only copy (pints/_util.py:77-95)."""
    if np.isscalar(x):
        x = np.array([float(x)])
    else:
        x = np.array(x, copy=True, dtype=float)
    if x.ndim != 1:
        n = np.max(x.shape)
        if np.prod(x.shape) != n:
            raise ValueError('Unable to convert to 1d vector of scalar values.')
        x = x.reshape((n,))
    x.setflags(write=False)
    return x


# --------------------------------------------------------------------------
# closed-form forward models
# --------------------------------------------------------------------------

def logistic_simulate(parameters, times, p0=2.0):
    """Logistic growth (pints/toy/_logistic_model.py:62-86).

    ``k / (1 + (k / p0 - 1) * exp(-r t))``; all zeros when ``p0 == 0`` or
    ``k < 0``; negative times are an error.
    """
    r, k = (float(v) for v in parameters)
    times = np.asarray(times)
    if np.any(times < 0):
        raise ValueError('Negative times are not allowed.')
    p0 = float(p0)
    if p0 == 0 or k < 0:
        return np.zeros(times.shape)
    decay = np.exp(-r * times)
    c = (k / p0 - 1)
    return k / (1 + c * decay)


def constant_simulate(parameters, times, force_multi_output=False):
    """Constant model: output o is ``p_o * (o + 1)`` at every time
    (pints/toy/_constant_model.py:75-93)."""
    parameters = np.asarray(parameters)
    times = np.asarray(times)
    n = len(parameters)
    if np.any(times < 0):
        raise ValueError('Negative times are not allowed.')
    if not np.all(np.isfinite(parameters)):
        raise ValueError('All parameters must be finite.')
    row = parameters.reshape((1, n)) * np.arange(1, 1 + n)
    out = row.repeat(len(times), axis=0)
    if n == 1 and not force_multi_output:
        out = out.reshape((len(times),))
    return out


def hh_ik_protocol():
    """Voltage-step protocol of the Hodgkin-Huxley IK toy model
    (pints/toy/_hh_ik_model.py:135-173).

    Returns ``(events, voltages, v_hold)``: the 24 times at which the clamp
    voltage changes, the voltage that applies *after* each event, and the
    voltage that applies before the first one.
    """
    t_hold, t_step = 90, 10
    period = t_hold + t_step
    v_hold = -(0 + 75)
    step_mv = [-6, -11, -19, -26, -32, -38, -51, -63, -76, -88, -100, -109]
    v_step = np.array([-(s + 75) for s in step_mv])
    n = len(v_step)
    events = np.concatenate((period * (1 + np.arange(n)),
                             period * np.arange(n) + t_hold))
    events.sort()
    voltages = np.repeat(v_step, 2)
    voltages[1::2] = v_hold
    return events, voltages, v_hold


def hh_ik_simulate(parameters, times, n0=0.3, g_max=36, e_k=-88):
    """Potassium current under the step protocol
    (pints/toy/_hh_ik_model.py:175-214).

    The gate ``n`` relaxes exponentially inside every constant-voltage segment;
    segments are half-open ``[t_last, t_next)`` so a sample on an event time
    belongs to the following segment, and everything at or after the final
    event uses the last voltage.  ``I = g_max * n**4 * (V - E_k)``.
    """
    if times[0] < 0:
        raise ValueError('All times must be positive.')
    times = np.asarray(times)
    p1, p2, p3, p4, p5 = parameters
    events, voltages, v_hold = hh_ik_protocol()

    def gate(v, n_start, t_start, t):
        a = p1 * (-(v + 75) + p2) / (np.exp((-(v + 75) + p2) / p3) - 1)
        b = p4 * np.exp((-v - 75) / p5)
        tau = 1 / (a + b)
        inf = a * tau
        return inf - (inf - n_start) * np.exp(-(t - t_start) / tau)

    ns = np.zeros(times.shape)
    vs = np.zeros(times.shape)
    v, t_last, n_last = v_hold, 0, float(n0)
    for i, t_next in enumerate(events):
        sel = (t_last <= times) * (times < t_next)
        vs[sel] = v
        ns[sel] = gate(v, n_last, t_last, times[sel])
        n_last = gate(v, n_last, t_last, t_next)
        t_last = t_next
        v = voltages[i]
    sel = times >= t_next
    vs[sel] = v
    ns[sel] = gate(v, n_last, t_last, times[sel])
    return g_max * ns**4 * (vs - e_k)


# --------------------------------------------------------------------------
# ODE forward models: LSODA through scipy.integrate.odeint, as the reference
# --------------------------------------------------------------------------

def fhn_rhs(y, t, p):
    """FitzHugh-Nagumo right-hand side
    (pints/toy/_fitzhugh_nagumo_model.py:111-117)."""
    V, R = y
    a, b, c = [float(v) for v in p]
    return (V - V**3 / 3 + R) * c, (V - a + b * R) / -c


def lv_rhs(state, t, p):
    """Lotka-Volterra right-hand side
    (pints/toy/_lotka_volterra_model.py:91-97)."""
    x, y = state
    a, b, c, d = p
    return np.array([a * x - b * x * y, -c * y + d * x * y])


def sir_rhs(y, t, gamma, v):
    """SIR right-hand side (pints/toy/_sir_model.py:96-103)."""
    dS = -gamma * y[0] * y[1]
    dI = gamma * y[0] * y[1] - v * y[1]
    dR = v * y[1]
    return np.array([dS, dI, dR])


def _toy_ode_simulate(rhs, y0, parameters, times, n_outputs, **tol):
    """Shared driver of the toy ODE models
    (pints/toy/_toy_classes.py:209-234): a leading ``t = 0`` is inserted when
    the first requested time is not zero, and dropped from the result."""
    times = as_vector(times)
    if np.any(times < 0):
        raise ValueError('Negative times are not allowed.')
    skip = 0
    if len(times) < 1 or times[0] != 0:
        times = np.concatenate(([0], times))
        skip = 1
    sol = odeint(rhs, y0, times, (parameters,), **tol)
    return sol[skip:, :n_outputs].squeeze()


FHN_Y0 = (-1.0, 1.0)
LV_Y0 = tuple(np.log([30, 4]))


def fhn_simulate(parameters, times, y0=None, **tol):
    """pints/toy/_fitzhugh_nagumo_model.py:68-77 (y0 default ``[-1, 1]``)."""
    y0 = np.array(FHN_Y0 if y0 is None else y0, dtype=float)
    return _toy_ode_simulate(fhn_rhs, y0, parameters, times, 2, **tol)


def lv_simulate(parameters, times, y0=None, **tol):
    """pints/toy/_lotka_volterra_model.py:45-49 (y0 default ``log([30, 4])``,
    kept as a Python list as the reference does, :138-145)."""
    y0 = list(LV_Y0) if y0 is None else list(y0)
    return _toy_ode_simulate(lv_rhs, y0, parameters, times, 2, **tol)


def sir_simulate(parameters, times, y0=None, **tol):
    """pints/toy/_sir_model.py:74-111.

    The initial condition applies at ``times[0]`` (no ``t = 0`` is inserted).
    With the default ``y0`` the state array is *integer*, so assigning ``S0``
    truncates it toward zero; a user ``y0`` is float and keeps ``S0`` as is.
    Output columns are (I, R).
    """
    gamma, v, S0 = parameters
    if y0 is None:
        start = np.array([38, 1, 0])             # int64, as the reference
    else:
        start = np.array(y0, dtype=float)
    start = np.array(start, copy=True)
    start[0] = S0
    sol = odeint(sir_rhs, start, times, (gamma, v), **tol)
    return sol[:, 1:]


def goodwin_rhs(state, time, parameters):
    """pints/toy/_goodwin_oscillator_model.py:101-107."""
    x, y, z = state
    k2, k3, m1, m2, m3 = parameters
    dxdt = 1 / (1 + z**10) - m1 * x
    dydt = k2 * x - m2 * y
    dzdt = k3 * y - m3 * z
    return dxdt, dydt, dzdt


GOODWIN_Y0 = (0.0054, 0.053, 1.93)


def goodwin_simulate(parameters, times, y0=None, **tol):
    """pints/toy/_goodwin_oscillator_model.py:19-20 (y0) through the shared
    driver (three states, all of them outputs)."""
    y0 = list(GOODWIN_Y0 if y0 is None else y0)
    return _toy_ode_simulate(goodwin_rhs, y0, parameters, times, 3, **tol)


def hes1_simulate(parameters, times, m0=2, fixed=(5., 3., 0.03), **tol):
    """pints/toy/_hes1_michaelis_menten.py:129-159: states (m, p1, p2), only m
    is returned; ``fixed = (p1_0, p2_0, k_deg)``."""
    kdeg = fixed[2]

    def rhs(state, time, parameters):
        m, p1, p2 = state
        P0, v, k1, h = parameters
        return np.array([
            - kdeg * m + 1. / (1. + (p2 / P0)**h),
            - kdeg * p1 + v * m - k1 * p1,
            - kdeg * p2 + k1 * p1])
    y0 = [m0, fixed[0], fixed[1]]
    return _toy_ode_simulate(rhs, y0, parameters, times, 1, **tol)


def repressilator_rhs(y, t, alpha_0, alpha, beta, n):
    """pints/toy/_repressilator_model.py:91-103."""
    dy = np.zeros(6)
    dy[0] = -y[0] + alpha / (1 + y[5]**n) + alpha_0
    dy[1] = -y[1] + alpha / (1 + y[3]**n) + alpha_0
    dy[2] = -y[2] + alpha / (1 + y[4]**n) + alpha_0
    dy[3] = -beta * (y[3] - y[0])
    dy[4] = -beta * (y[4] - y[1])
    dy[5] = -beta * (y[5] - y[2])
    return dy


def repressilator_simulate(parameters, times, y0=None, **tol):
    """pints/toy/_repressilator_model.py:105-108: integrated from ``times[0]``
    (no ``t = 0`` is inserted), the three mRNA states are returned; default
    y0 = [0, 0, 0, 2, 1, 3] (:68-70)."""
    alpha_0, alpha, beta, n = parameters
    start = np.array([0, 0, 0, 2, 1, 3]) if y0 is None \
        else np.array(y0, dtype=float)
    y = odeint(repressilator_rhs, start, times, (alpha_0, alpha, beta, n),
               **tol)
    return y[:, :3]


BEELER_REUTER_CONSTS = dict(
    y0=(-84.622, 2e-7), m0=0.01, h0=0.99, j0=0.98, d0=0.003, f0=0.99,
    x10=0.0004, C_m=1.0, E_Na=50.0, I_Stim_amp=25.0, I_Stim_period=1000.0,
    I_Stim_length=2.0)


def beeler_reuter_rhs(states, time, parameters, k=BEELER_REUTER_CONSTS):
    """pints/toy/_beeler_reuter_model.py:105-183 (Beeler-Reuter 1977): states
    (V, Ca_i, m, h, j, d, f, x1); the parameters are the logarithms of the
    maximum conductances; the stimulus is on while time % period < length."""
    V, Cai, m, h, j, d, f, x1 = states
    gNaBar, gNaC, gCaBar, gK1Bar, gx1Bar = np.exp(parameters)
    # INa
    INa = (gNaBar * m**3 * h * j + gNaC) * (V - k['E_Na'])
    alpha = (V + 47) / (1 - np.exp(-0.1 * (V + 47)))
    beta = 40 * np.exp(-0.056 * (V + 72))
    dmdt = alpha * (1 - m) - beta * m
    alpha = 0.126 * np.exp(-0.25 * (V + 77))
    beta = 1.7 / (1 + np.exp(-0.082 * (V + 22.5)))
    dhdt = alpha * (1 - h) - beta * h
    alpha = 0.055 * np.exp(-0.25 * (V + 78)) / (1 + np.exp(-0.2 * (V + 78)))
    beta = 0.3 / (1 + np.exp(-0.1 * (V + 32)))
    djdt = alpha * (1 - j) - beta * j
    # ICa
    E_Ca = -82.3 - 13.0287 * np.log(Cai)
    ICa = gCaBar * d * f * (V - E_Ca)
    alpha = 0.095 * np.exp(-0.01 * (V + -5)) / (np.exp(-0.072 * (V + -5)) + 1)
    beta = 0.07 * np.exp(-0.017 * (V + 44)) / (np.exp(0.05 * (V + 44)) + 1)
    dddt = alpha * (1 - d) - beta * d
    alpha = 0.012 * np.exp(-0.008 * (V + 28)) / (np.exp(0.15 * (V + 28)) + 1)
    beta = 0.0065 * np.exp(-0.02 * (V + 30)) / (np.exp(-0.2 * (V + 30)) + 1)
    dfdt = alpha * (1 - f) - beta * f
    # Cai
    dCaidt = -1e-7 * ICa + 0.07 * (1e-7 - Cai)
    # IK1
    IK1 = gK1Bar * (
        4 * (np.exp(0.04 * (V + 85)) - 1)
        / (np.exp(0.08 * (V + 53)) + np.exp(0.04 * (V + 53)))
        + 0.2 * (V + 23) / (1 - np.exp(-0.04 * (V + 23))))
    # IX1
    Ix1 = gx1Bar * x1 * (np.exp(0.04 * (V + 77)) - 1) \
        / np.exp(0.04 * (V + 35))
    alpha = 0.0005 * np.exp(0.083 * (V + 50)) / (np.exp(0.057 * (V + 50)) + 1)
    beta = 0.0013 * np.exp(-0.06 * (V + 20)) / (np.exp(-0.04 * (V + 333)) + 1)
    dx1dt = alpha * (1 - x1) - beta * x1
    # stimulus and membrane potential
    if (time % k['I_Stim_period']) < k['I_Stim_length']:
        IStim = k['I_Stim_amp']
    else:
        IStim = 0
    dVdt = -(1 / k['C_m']) * (IK1 + Ix1 + INa + ICa - IStim)
    return np.array([dVdt, dCaidt, dmdt, dhdt, djdt, dddt, dfdt, dx1dt])


def beeler_reuter_simulate(parameters, times, y0=None, all_states=False,
                           **tol):
    """pints/toy/_beeler_reuter_model.py:205-222: odeint from ``times[0]`` with
    ``hmax`` = the stimulus length and the model's own default tolerances
    rtol = 1e-4, atol = 1e-6 (:197-203); returns (V, Ca_i)."""
    k = BEELER_REUTER_CONSTS
    v0, cai0 = k['y0'] if y0 is None else y0
    start = [v0, cai0, k['m0'], k['h0'], k['j0'], k['d0'], k['f0'], k['x10']]
    opts = dict(rtol=1e-4, atol=1e-6)
    opts.update(tol)
    sol = odeint(beeler_reuter_rhs, start, times, args=(parameters,),
                 hmax=k['I_Stim_length'], **opts)
    return sol if all_states else sol[:, 0:2]


TRUTH_TOL = dict(rtol=1e-13, atol=1e-13, mxstep=100000)


# --------------------------------------------------------------------------
# model registry used by the spec-driven entry points below
# --------------------------------------------------------------------------

MODELS = ('logistic', 'hh_ik', 'constant', 'fhn', 'lotka_volterra', 'sir',
          'goodwin', 'hes1', 'repressilator', 'beeler_reuter')


def simulate(model, consts, parameters, times, **tol):
    """Dispatch on a model name; ``consts`` is the per-model constant dict
    (``p0`` / ``n0`` / ``y0`` / ``force_multi_output``).  Returns exactly what
    the reference's ``model.simulate(parameters, times)`` returns."""
    consts = dict(consts or {})
    if model == 'logistic':
        return logistic_simulate(parameters, times, consts.get('p0', 2.0))
    if model == 'hh_ik':
        return hh_ik_simulate(parameters, times, consts.get('n0', 0.3))
    if model == 'constant':
        return constant_simulate(parameters, times,
                                 consts.get('force_multi_output', False))
    if model == 'fhn':
        return fhn_simulate(parameters, times, consts.get('y0'), **tol)
    if model == 'lotka_volterra':
        return lv_simulate(parameters, times, consts.get('y0'), **tol)
    if model == 'sir':
        return sir_simulate(parameters, times, consts.get('y0'), **tol)
    if model == 'goodwin':
        return goodwin_simulate(parameters, times, consts.get('y0'), **tol)
    if model == 'hes1':
        return hes1_simulate(parameters, times, consts.get('m0', 2),
                             consts.get('fixed', (5., 3., 0.03)), **tol)
    if model == 'repressilator':
        return repressilator_simulate(parameters, times, consts.get('y0'),
                                      **tol)
    if model == 'beeler_reuter':
        return beeler_reuter_simulate(parameters, times, consts.get('y0'),
                                      **tol)
    raise ValueError('unknown model ' + str(model))


def problem_evaluate(model, consts, parameters, times, n_outputs,
                     multi_output, **tol):
    """``SingleOutputProblem.evaluate`` / ``MultiOutputProblem.evaluate``
    (pints/_core.py:147-153, :257-265): reshape to ``(n_t,)`` or
    ``(n_t, n_o)``."""
    y = np.asarray(simulate(model, consts, parameters, times, **tol))
    if multi_output:
        return y.reshape(len(times), n_outputs)
    return y.reshape((len(times),))


# --------------------------------------------------------------------------
# objectives
# --------------------------------------------------------------------------

def sum_of_squares_error(sim, values, weights):
    """pints/_error_measures.py:373-376."""
    return np.sum((np.sum(((sim - values)**2), axis=0) * weights), axis=0)


def mean_squared_error(sim, values, weights):
    """pints/_error_measures.py:100-113."""
    ninv = 1.0 / np.prod(values.shape)
    return ninv * np.sum(weights * np.sum((sim - values)**2, axis=0), axis=0)


def root_mean_squared_error(sim, values, normalised=False):
    """pints/_error_measures.py:201-229, and :135-161 when normalised
    (single-output problems)."""
    ninv = 1.0 / len(values)
    norm = 1.0 / np.sqrt(ninv * np.sum(values**2)) if normalised else None
    f = np.sqrt(ninv * np.sum((sim - values)**2))
    return norm * f if normalised else f


def gaussian_known_sigma_consts(n_times, n_outputs, sigma):
    """Pre-computed parts (pints/_log_likelihoods.py:1012-1035)."""
    if np.isscalar(sigma):
        sigma = np.ones(n_outputs) * float(sigma)
    else:
        sigma = as_vector(sigma)
        if len(sigma) != n_outputs:
            raise ValueError(
                'Sigma must be a scalar or a vector of length n_outputs.')
    if np.any(sigma <= 0):
        raise ValueError('Standard deviation must be greater than zero.')
    offset = -0.5 * n_times * np.log(2 * np.pi)
    offset = offset - n_times * np.log(sigma)
    multip = -1 / (2.0 * sigma**2)
    return offset, multip


def gaussian_known_sigma_ll(sim, values, offset, multip):
    """pints/_log_likelihoods.py:1039-1041 (error = data - simulation)."""
    error = values - sim
    return np.sum(offset + multip * np.sum(error**2, axis=0))


def gaussian_ll(sigma, sim, values, n_times):
    """pints/_log_likelihoods.py:1110-1129 after the sigma check."""
    logn = 0.5 * n_times * np.log(2 * np.pi)
    error = values - sim
    return np.sum(- logn - n_times * np.log(sigma)
                  - np.sum(error**2, axis=0) / (2 * sigma**2))


def composed_log_prior(parts, x):
    """``ComposedLogPrior.__call__`` (pints/_log_priors.py:212-219) over
    ``UniformLogPrior`` (:1315-1320, with ``RectangularBoundaries.check``,
    pints/_boundaries.py:151-157) and ``GaussianLogPrior`` (:463-468)."""
    output = 0
    lo = hi = 0
    for part in parts:
        if part[0] == 'uniform':
            lower, upper = np.asarray(part[1], float), np.asarray(part[2], float)
            lo, hi = hi, hi + len(lower)
            xx = x[lo:hi]
            inside = not (np.any(xx < lower) or np.any(xx >= upper))
            value = -np.log(np.prod(upper - lower))
            output += value if inside else -np.inf
        else:
            mean, sd = float(part[1]), float(part[2])
            lo, hi = hi, hi + 1
            offset = np.log(1 / np.sqrt(2 * np.pi * sd ** 2))
            factor = 1 / (2 * sd ** 2)
            output += offset - factor * (x[lo:hi][0] - mean)**2
    return output


class Spec(object):
    """One (model, series, objective) problem, described with plain data.

    model        one of MODELS
    consts       model constants (dict)
    times        (n_t,)
    values       (n_t,) for single-output problems, (n_t, n_o) for multi
    objective    'sse' | 'mse' | 'rmse' | 'nrmse' | 'gaussian_known_sigma' |
                 'gaussian'
    obj          'sse', 'mse': weights (n_o,); 'gaussian_known_sigma': sigma
    sign         +1, or -1 when wrapped in ProbabilityBasedError
                 (pints/_error_measures.py:183-184)
    prior_box    optional (lower, upper) of a UniformLogPrior combined through
                 LogPosterior (pints/_log_pdfs.py:207-212,
                 pints/_log_priors.py:1319-1320, pints/_boundaries.py:151-157)
    """

    def __init__(self, model, times, values, objective, obj=None, consts=None,
                 sign=1.0, prior_box=None, tol=None, prior_parts=None):
        self.model = model
        self.consts = dict(consts or {})
        self.times = as_vector(times)
        values = np.array(values, dtype=float)
        self.multi_output = values.ndim == 2
        self.values = values
        self.n_times = len(self.times)
        self.n_outputs = values.shape[1] if self.multi_output else 1
        self.objective = objective
        self.sign = sign
        self.prior_box = prior_box
        # ComposedLogPrior: [('uniform', lower, upper) | ('gaussian', mean, sd)]
        self.prior_parts = prior_parts
        self.tol = dict(tol or {})
        if objective in ('sse', 'mse'):
            w = [1] * self.n_outputs if obj is None else obj
            self.weights = np.asarray([float(v) for v in w])
        elif objective in ('rmse', 'nrmse'):
            if self.multi_output:
                raise ValueError('This measure is only defined for single '
                                 'output problems.')
        elif objective == 'gaussian_known_sigma':
            self.offset, self.multip = gaussian_known_sigma_consts(
                self.n_times, self.n_outputs, obj)
        elif objective != 'gaussian':
            raise ValueError('unknown objective ' + str(objective))

    def n_model_parameters(self):
        return {'logistic': 2, 'hh_ik': 5, 'fhn': 3, 'lotka_volterra': 4,
                'sir': 3, 'goodwin': 5, 'hes1': 4,
                'repressilator': 4,
                'beeler_reuter': 5}.get(self.model, self.n_outputs)

    def n_parameters(self):
        extra = self.n_outputs if self.objective == 'gaussian' else 0
        return self.n_model_parameters() + extra

    def trajectory(self, parameters, **tol):
        tol = tol or self.tol
        return problem_evaluate(self.model, self.consts, parameters,
                                self.times, self.n_outputs,
                                self.multi_output, **tol)

    def score(self, x, **tol):
        """f(x) exactly as the reference object graph would return it."""
        x = np.asarray(x, dtype=float)
        if self.prior_parts is not None:
            log_prior = composed_log_prior(self.prior_parts, x)
            # LogPosterior.__call__ (pints/_log_pdfs.py:207-212)
            if log_prior == -np.inf:
                return self.sign * -np.inf
        if self.prior_box is not None:
            lower, upper = self.prior_box
            # RectangularBoundaries.check: two negative tests, so a NaN
            # coordinate counts as inside (pints/_boundaries.py:151-157)
            outside = bool(np.any(x < lower) or np.any(x >= upper))
            if outside:
                return self.sign * -np.inf
            log_prior = -np.log(np.prod(np.asarray(upper) - np.asarray(lower)))
        if self.objective == 'gaussian':
            sigma = np.asarray(x[-self.n_outputs:])
            if any(sigma <= 0):
                f = -np.inf
            else:
                sim = self.trajectory(x[:-self.n_outputs], **tol)
                f = gaussian_ll(sigma, sim, self.values, self.n_times)
        else:
            sim = self.trajectory(x, **tol)
            if self.objective == 'sse':
                f = sum_of_squares_error(sim, self.values, self.weights)
            elif self.objective == 'mse':
                f = mean_squared_error(sim, self.values, self.weights)
            elif self.objective in ('rmse', 'nrmse'):
                f = root_mean_squared_error(sim, self.values,
                                            self.objective == 'nrmse')
            else:
                f = gaussian_known_sigma_ll(sim, self.values, self.offset,
                                            self.multip)
        if self.prior_box is not None or self.prior_parts is not None:
            f = log_prior + f
        return self.sign * f


def sequential_evaluate(function, positions, args=()):
    """``SequentialEvaluator._evaluate`` (pints/_evaluation.py:478-482),
    including the ``len()`` check of ``Evaluator.evaluate`` (:107-119)."""
    try:
        len(positions)
    except TypeError:
        raise ValueError(
            'The argument `positions` must be a sequence of input values'
            ' to the evaluator\'s function.')
    scores = [0] * len(positions)
    for k, x in enumerate(positions):
        scores[k] = function(x, *args)
    return scores


def scores(spec, xs, **tol):
    return np.array(sequential_evaluate(lambda x: spec.score(x, **tol), xs),
                    dtype=float)


# --------------------------------------------------------------------------
# samplers (ask/tell state machines, numpy legacy global RNG as the reference)
# --------------------------------------------------------------------------

def default_sigma0_single(x0):
    """pints/_mcmc/__init__.py:84-91."""
    s = np.abs(as_vector(x0))
    s = np.array(s)
    s[s == 0] = 1
    return np.diag(0.01 * s)


class HaarioBardenetChain(object):
    """One chain of HaarioBardenetACMC
    (pints/_mcmc/_adaptive_covariance.py:109-131, :236-303;
    pints/_mcmc/_haario_bardenet_ac.py:59-73).

    Draw order on numpy's global legacy stream: ``ask`` consumes the d normals
    of one ``multivariate_normal`` call; ``tell`` consumes one uniform, and only
    when the proposed log-pdf is finite.
    """

    def __init__(self, x0, sigma0=None, eta=0.6, target=0.234):
        self.x0 = as_vector(x0)
        self.d = len(self.x0)
        if sigma0 is None:
            self.sigma = default_sigma0_single(self.x0)
        else:
            s = np.array(sigma0, copy=True, dtype=float)
            self.sigma = np.diag(s.reshape(-1)) if s.size == self.d \
                else s.reshape((self.d, self.d))
        self.mu = np.array(self.x0, copy=True)
        self.eta = eta
        self.target = target
        self.gamma = 1
        self.adaptive = False
        self.adaptations = 1
        self.iterations = 0
        self.accepted_count = 0
        self.log_lambda = 0
        self.current = None
        self.current_f = None
        self.proposed = None
        self.started = False
        self.last_log_u = None        # recorded for replay tests

    def ask(self):
        if not self.started:
            self.started = True
            self.proposed = self.x0
        if self.proposed is None:
            self.proposed = np.random.multivariate_normal(
                self.current, self.sigma * np.exp(self.log_lambda))
        return self.proposed

    def tell(self, fx):
        fx = float(fx)
        self.iterations += 1
        self.last_log_u = None
        if self.current is None:
            if not np.isfinite(fx):
                raise ValueError(
                    'Initial point for MCMC must have finite logpdf.')
            self.current, self.current_f, self.proposed = \
                self.proposed, fx, None
            return self.current, self.current_f, True
        log_ratio = fx - self.current_f
        accepted = False
        if np.isfinite(fx):
            u = np.log(np.random.uniform(0, 1))
            self.last_log_u = u
            if u < log_ratio:
                accepted = True
                self.accepted_count += 1
                self.current = self.proposed
                self.current_f = fx
        self.proposed = None
        if self.adaptive:
            self.gamma = (self.adaptations + 1) ** -self.eta
            self.adaptations += 1
            self.mu = (1 - self.gamma) * self.mu + self.gamma * self.current
            dev = np.reshape(self.current - self.mu, (self.d, 1))
            self.sigma = ((1 - self.gamma) * self.sigma
                          + self.gamma * np.dot(dev, dev.T))
            p = 1 if accepted else 0
            self.log_lambda += self.gamma * (p - self.target)
        return self.current, self.current_f, accepted


class DifferentialEvolutionChains(object):
    """DifferentialEvolutionMCMC with its default settings
    (pints/_mcmc/_differential_evolution.py:60-124, :146-192, :283-332).

    Draw order per iteration: for every chain j, d normals then one
    ``choice(n - 1 candidates, 2, replace=False)``; in ``tell`` n uniforms.
    """

    def __init__(self, n_chains, x0):
        self.n = int(n_chains)
        if self.n < 3:
            raise ValueError('Need at least 3 chains.')
        self.x0 = np.array([as_vector(x) for x in x0])
        self.d = self.x0.shape[1]
        self.gamma_default = 2.38 / np.sqrt(2 * self.d)
        self.gamma = self.gamma_default
        self.switch_rate = 10
        self.b = 0.001
        self.current = None
        self.current_f = None
        self.proposed = None
        self.started = False
        self.last = {}                # eps, r1, r2, gamma, log_u for replay

    def ask(self):
        if not self.started:
            self.proposed = self.x0
            self.center = np.mean(self.x0, axis=0)
            self.b_star = np.abs(self.center * self.b)
            self.iter_count = 0
            self.started = True
        if self.proposed is None:
            if self.iter_count % self.switch_rate == 0:
                self.gamma = 1
            self.iter_count += 1
            self.proposed = np.zeros(self.current.shape)
            eps = np.zeros(self.current.shape)
            r1s = np.zeros(self.n, dtype=np.int64)
            r2s = np.zeros(self.n, dtype=np.int64)
            for j in range(self.n):
                eps[j] = np.random.normal(0, self.b_star, self.center.shape)
                others = list(range(self.n))
                others.pop(j)
                r1, r2 = np.random.choice(others, 2, replace=False)
                r1s[j], r2s[j] = r1, r2
                self.proposed[j] = (
                    self.current[j]
                    + self.gamma * (self.current[r1] - self.current[r2])
                    + eps[j])
            self.last = dict(eps=eps, r1=r1s, r2=r2s, gamma=float(self.gamma))
            self.gamma = self.gamma_default
        return self.proposed

    def tell(self, fxs):
        fxs = np.array(fxs)
        if self.current is None:
            if not np.all(np.isfinite(fxs)):
                raise ValueError(
                    'Initial points for MCMC must have finite logpdf.')
            self.current, self.current_f, self.proposed = \
                self.proposed, fxs, None
            return self.current, self.current_f, np.array([True] * self.n)
        nxt = np.array(self.current, copy=True)
        nxt_f = np.array(self.current_f, copy=True)
        u = np.log(np.random.uniform(size=self.n))
        self.last['log_u'] = u
        take = u < (fxs - self.current_f)
        nxt[take] = self.proposed[take]
        nxt_f[take] = fxs[take]
        self.current, self.current_f, self.proposed = nxt, nxt_f, None
        return self.current, self.current_f, take


# --------------------------------------------------------------------------
# benchmark problems: deterministic synthetic inputs
# --------------------------------------------------------------------------

def headline_fhn_spec(n_times=1000, sigma=0.5, seed=1):
    """FHN, 1000 points on [0, 20], Gaussian noise sd 0.5, known-sigma LL."""
    times = np.linspace(0, 20, n_times)
    rng = np.random.RandomState(seed)
    clean = fhn_simulate([0.1, 0.5, 3], times)
    values = clean + rng.normal(0, sigma, clean.shape)
    return Spec('fhn', times, values, 'gaussian_known_sigma', sigma)


def headline_fhn_points(n, seed=2, spread=0.05):
    rng = np.random.RandomState(seed)
    return np.array([0.1, 0.5, 3]) * (1 + spread * rng.randn(n, 3))
