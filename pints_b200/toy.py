"""
Mirror of the reference's toy time-series models on the hot path
(``pints/toy``) with batched, GPU-resident ``simulate``.

``model.simulate(parameters, times)`` keeps the reference's contract for one
parameter vector (same shapes, same errors); given ``(N, d)`` parameters it
returns ``(N, n_times)`` (single-output models) or ``(N, n_times, n_outputs)``.
Both go through ``pb200_simulate``; there is no CPU implementation here.
"""
import numpy as np

from . import _lib
from ._core import ForwardModel
from ._spec import ProblemSpec
from ._util import is_batch


class _DeviceSimulation(object):
    """Mixin: runs ``simulate`` on the GPU, caching one device problem per
    distinct time vector."""

    _sim_cache = None
    _sim_opts = None

    def set_solver_options(self, rtol=None, atol=None, max_steps=None,
                           device=None):
        """Integrator controls for the ODE models (ignored by closed forms);
        defaults equal scipy odeint's 1.49012e-8 used by the reference."""
        self._sim_opts = dict(rtol=rtol, atol=atol, max_steps=max_steps,
                              device=device)
        self._sim_cache = None

    def _device_problem(self, times):
        from ._engine import DeviceProblem      # needs torch + a GPU
        key = times.tobytes()
        cache = self._sim_cache
        if cache is None or cache[0] != key:
            from ._spec import describe_model
            model_id, consts = describe_model(self)
            n_o = self.n_outputs()
            spec = ProblemSpec(model_id, consts, times,
                               np.zeros((len(times), n_o)), _lib.OBJ_SSE,
                               [1.0] * n_o)
            opts = dict(self._sim_opts or {})
            device = opts.pop('device', None)
            cache = (key, DeviceProblem(spec, device, **opts))
            self._sim_cache = cache
        return cache[1]

    def _simulate_device(self, parameters, times):
        import torch
        times = np.ascontiguousarray(times, dtype=np.float64).reshape(-1)
        batch = is_batch(parameters)
        xs = np.asarray(parameters, dtype=np.float64)
        xs = xs if batch else xs.reshape((1, -1))
        d = self.n_parameters()
        if xs.shape[1] != d:
            raise ValueError('Expected %d parameters, got %d.'
                             % (d, xs.shape[1]))
        n_o = self.n_outputs()
        if len(times) == 0 or xs.shape[0] == 0:
            out = np.zeros((xs.shape[0], len(times), n_o))
        else:
            prob = self._device_problem(times)
            d_x = torch.from_numpy(np.ascontiguousarray(xs)).to(prob.device)
            out = prob.simulate(d_x).cpu().numpy()
        return out, batch


class LogisticModel(_DeviceSimulation, ForwardModel):
    """Logistic growth ``k / (1 + (k/p0 - 1) exp(-r t))``; parameters
    ``(r, k)`` (pints/toy/_logistic_model.py:14-96)."""

    def __init__(self, initial_population_size=2):
        super().__init__()
        self._p0 = float(initial_population_size)
        if self._p0 < 0:
            raise ValueError('Population size cannot be negative.')

    def n_parameters(self):
        return 2

    def simulate(self, parameters, times):
        times = np.asarray(times)
        if np.any(times < 0):
            raise ValueError('Negative times are not allowed.')
        out, batch = self._simulate_device(parameters, times)
        out = out[:, :, 0]
        return out if batch else out[0].reshape(times.shape)

    def suggested_parameters(self):
        return np.array([0.1, 50])

    def suggested_times(self):
        return np.linspace(0, 100, 100)


class ConstantModel(_DeviceSimulation, ForwardModel):
    """Output ``o`` equals ``p_o * (o + 1)`` at all times
    (pints/toy/_constant_model.py:13-113)."""

    def __init__(self, n, force_multi_output=False):
        super().__init__()
        n = int(n)
        if n < 1:
            raise ValueError('Number of parameters must be 1 or greater.')
        self._r = np.arange(1, 1 + n)
        self._n = n
        self._reshape = (n == 1 and not force_multi_output)

    def n_outputs(self):
        return self._n

    def n_parameters(self):
        return self._n

    def simulate(self, parameters, times):
        parameters = np.asarray(parameters)
        times = np.asarray(times)
        if np.any(times < 0):
            raise ValueError('Negative times are not allowed.')
        if parameters.shape[-1] != self._n:
            raise ValueError('Expected ' + str(self._n) + ' parameters.')
        if not np.all(np.isfinite(parameters)):
            raise ValueError('All parameters must be finite.')
        out, batch = self._simulate_device(parameters, times)
        if self._reshape:
            out = out[:, :, 0]
        return out if batch else out[0]

    def suggested_parameters(self):
        return self._r * 2.0

    def suggested_times(self):
        return np.linspace(0, 1, 100)


class HodgkinHuxleyIKModel(_DeviceSimulation, ForwardModel):
    """Potassium current of the Hodgkin-Huxley squid axon model under a
    12-step voltage clamp protocol; 5 parameters
    (pints/toy/_hh_ik_model.py:14-241).  Closed form, not an ODE solve."""

    def __init__(self, initial_condition=0.3):
        super().__init__()
        self._n0 = float(initial_condition)
        if self._n0 <= 0 or self._n0 >= 1:
            raise ValueError('Initial condition must be > 0 and < 1.')
        self._E_k = -88
        self._g_max = 36
        self._prepare_protocol()

    def _prepare_protocol(self):
        # 12 sweeps of 90 ms at the holding potential then 10 ms at a step
        # potential (pints/toy/_hh_ik_model.py:135-173)
        self._t_hold, self._t_step = 90, 10
        self._t_both = self._t_hold + self._t_step
        self._v_hold = -(0 + 75)
        depol = [-6, -11, -19, -26, -32, -38, -51, -63, -76, -88, -100, -109]
        self._v_step = np.array([-(m + 75) for m in depol])
        self._n_steps = len(self._v_step)
        self._duration = self._n_steps * self._t_both
        sweep = np.arange(self._n_steps)
        ends = self._t_both * (1 + sweep)
        step_on = self._t_both * sweep + self._t_hold
        self._events = np.sort(np.concatenate((ends, step_on)))
        self._voltages = np.repeat(self._v_step, 2)
        self._voltages[1::2] = self._v_hold

    def n_parameters(self):
        return 5

    def simulate(self, parameters, times):
        if times[0] < 0:
            raise ValueError('All times must be positive.')
        times = np.asarray(times)
        out, batch = self._simulate_device(parameters, times)
        out = out[:, :, 0]
        return out if batch else out[0].reshape(times.shape)

    def fold(self, times, values):
        """Splits a trace into one ``(times, values)`` pair per voltage step
        (pints/toy/_hh_ik_model.py:98-128)."""
        times = np.mod(times, self._t_both)
        keep = times >= self._t_hold
        times, values = times[keep] - self._t_hold, values[keep]
        cuts = 1 + np.argwhere(times[1:] < times[:-1]).reshape(-1)
        return list(zip(np.split(times, cuts), np.split(values, cuts)))

    def suggested_duration(self):
        return self._duration

    def suggested_parameters(self):
        return 0.01, 10, 10, 0.125, 80

    def suggested_times(self):
        fs = 4
        return np.arange(self._duration * fs) / fs


class _ToyODE(_DeviceSimulation, ForwardModel):
    """Common part of the ODE toy models (pints/toy/_toy_classes.py:59-240):
    integration starts at ``t = 0`` from ``y0`` whatever ``times[0]`` is."""

    def initial_conditions(self):
        return self._y0

    def n_states(self):
        return self.n_outputs()

    def simulate(self, parameters, times):
        times = np.asarray(times, dtype=float).reshape(-1)
        if np.any(times < 0):
            raise ValueError('Negative times are not allowed.')
        out, batch = self._simulate_device(parameters, times)
        return out if batch else out[0].squeeze()


class FitzhughNagumoModel(_ToyODE):
    """FitzHugh-Nagumo: ``V' = c (V - V^3/3 + R)``, ``R' = -(V - a + b R)/c``;
    parameters ``(a, b, c)``, outputs ``(V, R)``
    (pints/toy/_fitzhugh_nagumo_model.py:14-125)."""

    def __init__(self, y0=None):
        super().__init__()
        if y0 is None:
            self._y0 = np.array([-1, 1], dtype=float)
        else:
            self._y0 = np.array(y0, dtype=float)
            if len(self._y0) != 2:
                raise ValueError('Initial value must have size 2.')

    def set_initial_conditions(self, y0):
        self._y0 = y0
        self._sim_cache = None

    def n_outputs(self):
        return 2

    def n_parameters(self):
        return 3

    def suggested_parameters(self):
        return np.array([0.1, 0.5, 3])

    def suggested_times(self):
        return np.linspace(0, 20, 200)


class LotkaVolterraModel(_ToyODE):
    """Lotka-Volterra: ``x' = a x - b x y``, ``y' = -c y + d x y``; parameters
    ``(a, b, c, d)`` (pints/toy/_lotka_volterra_model.py:13-145)."""

    def __init__(self, y0=None):
        super().__init__()
        if y0 is None:
            self.set_initial_conditions(np.log([30, 4]))
        else:
            self.set_initial_conditions(y0)

    def initial_conditions(self):
        return np.array(self._y0, copy=True)

    def set_initial_conditions(self, y0):
        a, b = y0
        if a < 0 or b < 0:
            raise ValueError('Initial populations cannot be negative.')
        self._y0 = [a, b]
        self._sim_cache = None

    def n_outputs(self):
        return 2

    def n_parameters(self):
        return 4

    def suggested_parameters(self):
        return np.array([0.5, 0.15, 1.0, 0.3])

    def suggested_times(self):
        return np.linspace(0, 20, 21)

    def suggested_values(self):
        """Hudson's Bay Company hare / lynx pelt counts 1900-1920 (thousands),
        the data set the reference ships
        (pints/toy/_lotka_volterra_model.py:115-145)."""
        hare = [30.0, 47.2, 70.2, 77.4, 36.3, 20.6, 18.1, 21.4, 22.0, 25.4,
                27.1, 40.3, 57.0, 76.6, 52.3, 19.5, 11.2, 7.6, 14.6, 16.2,
                24.7]
        lynx = [4.0, 6.1, 9.8, 35.2, 59.4, 41.7, 19.0, 13.0, 8.3, 9.1, 7.4,
                8.0, 12.3, 19.5, 45.7, 51.1, 29.7, 15.8, 9.7, 10.1, 8.1]
        return np.column_stack((hare, lynx))


class SIRModel(_ToyODE):
    """SIR epidemic: ``S' = -g S I``, ``I' = g S I - v I``, ``R' = v I``;
    parameters ``(gamma, v, S0)``, outputs ``(I, R)``; the initial state holds
    at ``times[0]`` (pints/toy/_sir_model.py:14-155).  With the default
    (integer) ``y0`` the reference truncates ``S0`` to an integer; this mirror
    keeps that behaviour."""

    def __init__(self, y0=None):
        super().__init__()
        if y0 is None:
            self._y0 = np.array([38, 1, 0])
        else:
            self._y0 = np.array(y0, dtype=float)
            if len(self._y0) != 3:
                raise ValueError('Initial value must have size 3.')
            if np.any(self._y0 < 0):
                raise ValueError('Initial states can not be negative.')

    def n_outputs(self):
        return 2

    def n_parameters(self):
        return 3

    def n_states(self):
        return 3

    def simulate(self, parameters, times):
        # the reference's SIR does not validate times (it hands them to odeint
        # as they are), the mirror needs them non-decreasing like a problem does
        times = np.asarray(times, dtype=float).reshape(-1)
        out, batch = self._simulate_device(parameters, times)
        return out if batch else out[0]

    def suggested_parameters(self):
        return [0.026, 0.285, 38]

    def suggested_times(self):
        return np.arange(1, 22)

    def suggested_values(self):
        """Common-cold outbreak on Tristan da Cunha, days 1-21, (I, R) counts
        (pints/toy/_sir_model.py:126-155)."""
        infected = [1, 1, 3, 7, 6, 10, 13, 13, 14, 14, 17, 10, 6, 6, 4, 3, 1,
                    1, 1, 1, 0]
        recovered = [0, 0, 0, 0, 5, 7, 8, 13, 13, 16, 17, 24, 30, 31, 33, 34,
                     36, 36, 36, 36, 37]
        return np.column_stack((infected, recovered))


class GoodwinOscillatorModel(_ToyODE):
    """Goodwin oscillator: ``x' = 1/(1 + z^10) - m1 x``, ``y' = k2 x - m2 y``,
    ``z' = k3 y - m3 z``; parameters ``(k2, k3, m1, m2, m3)``, three outputs
    (pints/toy/_goodwin_oscillator_model.py:14-115)."""

    def __init__(self):
        super().__init__()
        self._y0 = [0.0054, 0.053, 1.93]

    def set_initial_conditions(self, y0):
        self._y0 = y0
        self._sim_cache = None

    def n_outputs(self):
        return 3

    def n_parameters(self):
        return 5

    def suggested_parameters(self):
        return np.array([2, 4, 0.12, 0.08, 0.1])

    def suggested_times(self):
        return np.linspace(0, 100, 200)


class Hes1Model(_ToyODE):
    """HES1 Michaelis-Menten model: states ``(m, p1, p2)``, only ``m`` is
    observed; parameters ``(P0, v, k1, h)``; ``m0`` and the fixed parameters
    ``(p1_0, p2_0, k_deg)`` as in the reference
    (pints/toy/_hes1_michaelis_menten.py:14-190)."""

    def __init__(self, m0=None, fixed_parameters=None):
        super().__init__()
        self.set_fixed_parameters([5., 3., 0.03] if fixed_parameters is None
                                  else fixed_parameters)
        self.set_m0(2 if m0 is None else m0)

    def m0(self):
        return self._y0[0]

    def fixed_parameters(self):
        return [self._p0[0], self._p0[1], self._kdeg]

    def set_m0(self, m0):
        if m0 < 0:
            raise ValueError('Initial condition cannot be negative.')
        self._y0 = [m0, self._p0[0], self._p0[1]]
        self._sim_cache = None

    def set_fixed_parameters(self, k):
        a, b, c = k
        if a < 0 or b < 0 or c < 0:
            raise ValueError('Implicit parameters cannot be negative.')
        self._p0 = [a, b]
        self._kdeg = c
        self._sim_cache = None

    def n_states(self):
        return 3

    def n_outputs(self):
        return 1

    def n_parameters(self):
        return 4

    def suggested_parameters(self):
        return np.array([2.4, 0.025, 0.11, 6.9])

    def suggested_times(self):
        return np.arange(0, 270, 30)

    def suggested_values(self):
        return np.array([2, 1.20, 5.90, 4.58, 2.64, 5.38, 6.42, 5.60, 4.48])


class RepressilatorModel(_ToyODE):
    """Repressilator: three mRNA and three protein levels, the mRNA levels are
    the outputs; parameters ``(alpha_0, alpha, beta, n)``; like ``SIRModel`` it
    is integrated from ``times[0]`` (pints/toy/_repressilator_model.py:14-125).
    """

    def __init__(self, y0=None):
        super().__init__()
        if y0 is None:
            self._y0 = np.array([0, 0, 0, 2, 1, 3])
        else:
            self._y0 = np.array(y0, dtype=float)
            if len(self._y0) != 6:
                raise ValueError('Initial value must have size 6.')
            if np.any(self._y0 < 0):
                raise ValueError('Initial states can not be negative.')

    def n_states(self):
        return 6

    def n_outputs(self):
        return 3

    def n_parameters(self):
        return 4

    def simulate(self, parameters, times):
        times = np.asarray(times, dtype=float).reshape(-1)
        out, batch = self._simulate_device(parameters, times)
        return out if batch else out[0]

    def suggested_parameters(self):
        return np.array([1, 1000, 5, 2])

    def suggested_times(self):
        return np.linspace(0, 40, 400)


class ActionPotentialModel(_DeviceSimulation, ForwardModel):
    """Beeler-Reuter (1977) ventricular action potential: eight states, the
    outputs are the membrane potential and the calcium transient; the five
    parameters are the natural logarithms of the maximum conductances
    (pints/toy/_beeler_reuter_model.py:16-229).  Like ``SIRModel`` it is
    integrated from ``times[0]``.

    The model is stiff (the reference's LSODA switches to BDF on it), so it
    does not run on the explicit DP5(4) kernel of the other ODE models but on a
    4-stage Rosenbrock method (``br_rosenbrock_kernel``).
    ``set_solver_tolerances`` keeps the reference's signature; the default is
    1e-7 for both rather than the reference's (1e-4, 1e-6): about 1 650 steps
    for 400 ms and 8.5e-7 from the converged log-likelihood, where the
    reference's default run is 1.7e-3 away.  The absolute tolerance of the
    calcium concentration (values of 1e-7 .. 1e-5) is ``atol * 1e-6``.
    """

    def __init__(self, y0=None):
        super().__init__()
        if y0 is None:
            self.set_initial_conditions([-84.622, 2e-7])
        else:
            self.set_initial_conditions(y0)
        # non-observable states, membrane capacitance, reversal potential and
        # stimulus (amplitude, period, length): fixed, as in the reference
        self._m0, self._h0, self._j0 = 0.01, 0.99, 0.98
        self._d0, self._f0, self._x10 = 0.003, 0.99, 0.0004
        self._C_m = 1.0
        self._E_Na = 50.0
        self._I_Stim_amp = 25.0
        self._I_Stim_period = 1000.0
        self._I_Stim_length = 2.0

    def initial_conditions(self):
        return [self._v0, self._cai0]

    def set_initial_conditions(self, y0):
        if y0[1] < 0:
            raise ValueError('Initial condition of ``cai`` cannot be'
                             ' negative.')
        self._v0 = y0[0]
        self._cai0 = y0[1]
        self._sim_cache = None

    def set_solver_tolerances(self, rtol=1e-7, atol=1e-7):
        self.set_solver_options(rtol=float(rtol), atol=float(atol))

    def n_outputs(self):
        return 2

    def n_parameters(self):
        return 5

    def n_states(self):
        return 8

    def simulate(self, parameters, times):
        times = np.asarray(times, dtype=float).reshape(-1)
        out, batch = self._simulate_device(parameters, times)
        return out if batch else out[0]

    def suggested_parameters(self):
        return np.log([4.0, 0.003, 0.09, 0.35, 0.8])

    def suggested_times(self):
        return np.arange(0, 400, 0.5)
