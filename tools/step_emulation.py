"""NumPy emulation of the device's DP5(4) step sequence on the headline FHN
population (CPU only): per-vector attempted steps and, per attempt, how many
observations fall into the accepted step.  Used to cost warp-level scheduling
policies (sort keys, in-line interpolation slots) without a GPU:

    python tools/step_emulation.py [n_vectors]

The emulation follows ode_kernel in pb200_ode.cu (error scale atol + rtol
(|y|+|y_new|)/2, mean-square norm, 0.9 err^-1/5 in [0.2, 10], no growth after a
rejection, Hairer's first step); it is not bit-exact (fp32 step factor, FMA
contraction), the step counts agree to a fraction of a step on average.
"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pints_oracle as po      # tools may use the oracle (inputs only)

A = [[], [1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
     [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
     [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
     [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
E = [71 / 57600, 0, -71 / 16695, 71 / 1920, -17253 / 339200, 22 / 525, -1 / 40]


def emulate(xs, times, rtol=1.49012e-8, atol=1.49012e-8, max_attempts=4000):
    """Returns (attempts (n,), pending (max_attempts, n) int8): pending[k, i]
    = observations inside the step vector i accepted at its k-th loop trip
    (0 when rejected or finished)."""
    n = xs.shape[0]
    a, b, c = xs[:, 0], xs[:, 1], xs[:, 2]

    def f(y):
        V, R = y
        return np.array([(V - V ** 3 / 3 + R) * c, (V - a + b * R) / -c])

    y = np.array([np.full(n, -1.0), np.full(n, 1.0)])
    t = np.zeros(n)
    k1 = f(y)
    sc = atol + rtol * np.abs(y)
    d0 = np.sqrt(np.mean((y / sc) ** 2, axis=0))
    d1 = np.sqrt(np.mean((k1 / sc) ** 2, axis=0))
    h_a = np.where((d0 < 1e-5) | (d1 < 1e-5), 1e-6, 0.01 * d0 / d1)
    k2 = f(y + h_a * k1)
    d2 = np.sqrt(np.mean(((k2 - k1) / sc) ** 2, axis=0)) / h_a
    dm = np.maximum(d1, d2)
    h = np.minimum(100 * h_a, np.where(dm <= 1e-15, np.maximum(1e-6, h_a * 1e-3),
                                       (0.01 / dm) ** 0.2))
    t_last = times[-1]
    nobs = np.searchsorted(times, t, side='right')      # observations consumed
    attempts = np.zeros(n, dtype=np.int32)
    after_reject = np.zeros(n, dtype=bool)
    pending = []
    while True:
        fin = nobs >= len(times)
        if fin.all() or len(pending) >= max_attempts:
            break
        working = ~fin & (t < t_last)
        attempts += working
        ks = [k1]
        for s in range(1, 7):
            w = y + h * sum(A[s][j] * ks[j] for j in range(s) if A[s][j] != 0)
            ks.append(f(w))
        yn = w
        err = h * sum(E[j] * ks[j] for j in range(7) if E[j] != 0)
        scale = atol + 0.5 * rtol * (np.abs(y) + np.abs(yn))
        sum2 = np.sum((err / scale) ** 2, axis=0)
        ok = ~fin & (sum2 <= 2.0)
        with np.errstate(divide='ignore'):
            fac = np.clip(0.9 * (sum2 / 2.0) ** -0.1, 0.2, 10.0)
        fac = np.where(~ok | after_reject, np.minimum(fac, 1.0), fac)
        after_reject = ~ok
        t_new = t + h
        new_nobs = np.where(ok, np.searchsorted(times, t_new, side='right'), nobs)
        pending.append((new_nobs - nobs).astype(np.int16))
        nobs = new_nobs
        t = np.where(ok, t_new, t)
        y = np.where(ok, yn, y)
        k1 = np.where(ok, ks[6], k1)
        h = h * fac
    return attempts, np.array(pending)


def warp_cost(attempts, pending, order, U=2, base=348.0, per_pass=37.0):
    """Issue cycles per vector when the vectors run 32 per warp in `order`:
    a warp executes max-over-lanes attempts and, per attempt, U in-line
    interpolation passes plus max(0, max-over-lanes pending of the PREVIOUS
    trip - U) loop passes."""
    n = len(order) // 32 * 32
    o = order[:n].reshape(-1, 32)
    wa = attempts[o].max(axis=1)                      # loop trips per warp
    trips = pending.shape[0]
    pw = pending[:, o].max(axis=2)                    # (trips, warps)
    alive = np.arange(trips)[:, None] < wa[None, :] + 1
    passes = np.where(alive, np.maximum(pw, U), 0).sum(axis=0)
    cyc = wa * base + passes * per_pass
    return dict(mean_attempts=float(attempts[order[:n]].mean()),
                warp_attempts=float(wa.mean()),
                passes_per_attempt=float(passes.sum() / wa.sum()),
                needed_per_attempt=float(pending[:, order[:n]].sum() / attempts[order[:n]].sum()),
                cycles_per_vector_attempt=float(cyc.sum() / attempts[order[:n]].sum() / 1.0))


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    spec = po.headline_fhn_spec()
    xs = po.headline_fhn_points(65536)[:n]
    att, pend = emulate(xs, spec.times)
    print('vectors %d  mean attempts %.2f  min %d  max %d  loop trips %d'
          % (n, att.mean(), att.min(), att.max(), pend.shape[0]))
    rng = np.random.RandomState(0)
    orders = {
        'unsorted': np.arange(n),
        'sorted by c (1024 bins of the whole population ~ %d here)' % max(1, n // 64):
            np.argsort(np.floor((xs[:, 2] - xs[:, 2].min()) / np.ptp(xs[:, 2]) * 1024 * n / 65536), kind='stable'),
        'sorted by c exactly': np.argsort(xs[:, 2], kind='stable'),
        'sorted by attempts (oracle)': np.argsort(att, kind='stable'),
    }
    # quadratic regression of the attempts on (a, b, c)
    X = xs / xs.mean(axis=0) - 1
    cols = [np.ones(n)] + [X[:, i] for i in range(3)] + \
           [X[:, i] * X[:, j] for i in range(3) for j in range(i, 3)]
    Q = np.array(cols).T
    coef, *_ = np.linalg.lstsq(Q, att, rcond=None)
    pred = Q @ coef
    print('quadratic fit of attempts: residual sd %.2f (attempts sd %.2f)'
          % ((att - pred).std(), att.std()))
    orders['sorted by quadratic fit of attempts'] = np.argsort(pred, kind='stable')
    for U in (2,):
        for name, order in orders.items():
            r = warp_cost(att, pend, order, U=U)
            print('U=%d %-55s warp attempts %.1f (lanes %.1f)  passes/attempt %.2f (needed %.2f)  '
                  'cycles per lane-attempt %.1f'
                  % (U, name, r['warp_attempts'], r['mean_attempts'], r['passes_per_attempt'],
                     r['needed_per_attempt'], r['cycles_per_vector_attempt']))
