"""NumPy prototype of the stiff integrator for the Beeler-Reuter model
(development tool; CPU only): the 4-stage, order-4(3) Rosenbrock method of Kaps
& Rentrop with Shampine's parameters, the Jacobian taken analytically where
the states enter polynomially and by ONE forward difference for the voltage
column, steps landing on the stimulus edges, cubic Hermite dense output.
Vectorised over parameter vectors; compares against the converged solutions in
tests/golden/beeler_reuter.npz and counts steps.

    python tools/br_rosenbrock_proto.py
"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pints_oracle as po          # inputs / checker only

K = po.BEELER_REUTER_CONSTS
GAM = 0.5
A21, A31, A32 = 2.0, 48.0 / 25.0, 6.0 / 25.0
C21, C31, C32 = -8.0, 372.0 / 25.0, 12.0 / 5.0
C41, C42, C43 = -112.0 / 125.0, -54.0 / 125.0, -2.0 / 5.0
B1, B2, B3, B4 = 19.0 / 9.0, 0.5, 25.0 / 108.0, 125.0 / 108.0
E1, E2, E3, E4 = 17.0 / 54.0, 7.0 / 36.0, 0.0, 125.0 / 108.0


def rhs(y, g, stim):
    """y (8, n); g (5, n) conductances; stim (n,) current.  Returns f (8, n) and
    the pieces the analytic Jacobian columns need."""
    V, Cai, m, h, j, d, f, x1 = y
    gNa, gNaC, gCa, gK1, gx1 = g
    e = np.exp
    am = (V + 47) / (1 - e(-0.1 * (V + 47))); bm = 40 * e(-0.056 * (V + 72))
    ah = 0.126 * e(-0.25 * (V + 77)); bh = 1.7 / (1 + e(-0.082 * (V + 22.5)))
    aj = 0.055 * e(-0.25 * (V + 78)) / (1 + e(-0.2 * (V + 78))); bj = 0.3 / (1 + e(-0.1 * (V + 32)))
    ad = 0.095 * e(-0.01 * (V - 5)) / (e(-0.072 * (V - 5)) + 1); bd = 0.07 * e(-0.017 * (V + 44)) / (e(0.05 * (V + 44)) + 1)
    af = 0.012 * e(-0.008 * (V + 28)) / (e(0.15 * (V + 28)) + 1); bf = 0.0065 * e(-0.02 * (V + 30)) / (e(-0.2 * (V + 30)) + 1)
    ax = 0.0005 * e(0.083 * (V + 50)) / (e(0.057 * (V + 50)) + 1); bx = 0.0013 * e(-0.06 * (V + 20)) / (e(-0.04 * (V + 333)) + 1)
    ECa = -82.3 - 13.0287 * np.log(Cai)
    m3 = m * m * m
    INa = (gNa * m3 * h * j + gNaC) * (V - K['E_Na'])
    ICa = gCa * d * f * (V - ECa)
    IK1 = gK1 * (4 * (e(0.04 * (V + 85)) - 1) / (e(0.08 * (V + 53)) + e(0.04 * (V + 53)))
                 + 0.2 * (V + 23) / (1 - e(-0.04 * (V + 23))))
    xf = (e(0.04 * (V + 77)) - 1) / e(0.04 * (V + 35))
    Ix1 = gx1 * x1 * xf
    out = np.array([-(IK1 + Ix1 + INa + ICa - stim) / K['C_m'],
                    -1e-7 * ICa + 0.07 * (1e-7 - Cai),
                    am * (1 - m) - bm * m, ah * (1 - h) - bh * h, aj * (1 - j) - bj * j,
                    ad * (1 - d) - bd * d, af * (1 - f) - bf * f, ax * (1 - x1) - bx * x1])
    diag = -np.array([am + bm, ah + bh, aj + bj, ad + bd, af + bf, ax + bx])   # d gate' / d gate
    return out, dict(diag=diag, ECa=ECa, m3=m3, xf=xf)


def jacobian(y, g, stim, f0, aux):
    """Arrow-shaped Jacobian: column V by a forward difference (8 entries), rows
    V and Cai analytically in (Cai, m, h, j, d, f, x1), the gates' diagonal."""
    V, Cai, m, h, j, d, f, x1 = y
    gNa, gNaC, gCa, gK1, gx1 = g
    dV = 1e-7 * np.maximum(np.abs(V), 1.0) * 10
    y2 = y.copy(); y2[0] = V + dV
    f1, _ = rhs(y2, g, stim)
    colV = (f1 - f0) / dV                                    # (8, n)
    dn = V - K['E_Na']
    icm = 1.0 / K['C_m']
    dECa = -13.0287 / Cai
    rowV = dict(Cai=-icm * gCa * d * f * (-dECa),
                m=-icm * gNa * 3 * m * m * h * j * dn, h=-icm * gNa * aux['m3'] * j * dn,
                j=-icm * gNa * aux['m3'] * h * dn, d=-icm * gCa * f * (V - aux['ECa']),
                f=-icm * gCa * d * (V - aux['ECa']), x1=-icm * gx1 * aux['xf'])
    rowC = dict(Cai=-1e-7 * gCa * d * f * (-dECa) - 0.07,
                d=-1e-7 * gCa * f * (V - aux['ECa']), f=-1e-7 * gCa * d * (V - aux['ECa']))
    return colV, rowV, rowC, aux['diag']


def solve(a, colV, rowV, rowC, diag, b):
    """(a I - J) x = b for the arrow-shaped J: gates eliminated first, then the
    2 x 2 system in (V, Cai)."""
    pg = 1.0 / (a - diag)                                    # (6, n)
    names = ['m', 'h', 'j', 'd', 'f', 'x1']
    # gate_i = pg_i (b_i + colV_i xV)
    aVV = a - colV[0]; aVC = -rowV['Cai']; bV = b[0].copy()
    aCV = -colV[1].copy(); aCC = a - rowC['Cai']; bC = b[1].copy()
    for i, nm in enumerate(names):
        r = rowV[nm] * pg[i]
        aVV -= r * colV[2 + i]; bV += r * b[2 + i]
        if nm in rowC:
            r = rowC[nm] * pg[i]
            aCV -= r * colV[2 + i]; bC += r * b[2 + i]
    det = aVV * aCC - aVC * aCV
    xV = (bV * aCC - aVC * bC) / det
    xC = (aVV * bC - aCV * bV) / det
    x = np.empty_like(b)
    x[0], x[1] = xV, xC
    for i in range(6):
        x[2 + i] = pg[i] * (b[2 + i] + colV[2 + i] * xV)
    return x


def integrate(xs, times, rtol, atol, weights=(1, 1e-6, 1, 1, 1, 1, 1, 1), max_attempts=20000):
    n = xs.shape[0]
    g = np.exp(xs[:, :5].T)
    y = np.array([np.full(n, v) for v in (K['y0'][0], K['y0'][1], K['m0'], K['h0'], K['j0'],
                                          K['d0'], K['f0'], K['x10'])])
    w = np.array(weights, dtype=float)[:, None]
    period, length, amp = K['I_Stim_period'], K['I_Stim_length'], K['I_Stim_amp']
    t = np.full(n, times[0])
    h = np.full(n, 1e-3)
    out = np.zeros((n, len(times), 2))
    nobs = np.zeros(n, dtype=int)
    attempts = np.zeros(n, dtype=int)
    accepted = np.zeros(n, dtype=int)

    def next_edge(t):
        ph = np.mod(t, period)
        return np.where(ph < length, t - ph + length, t - ph + period)

    # observations at the initial time
    first = times <= times[0]
    out[:, first, :] = y[:2].T[:, None, :]
    nobs[:] = first.sum()
    t_last = times[-1]
    for it in range(max_attempts):
        active = nobs < len(times)
        if not active.any():
            break
        attempts += active
        edge = next_edge(t)
        hs = np.minimum(h, edge - t)
        clipped = hs < h
        stim = np.where(np.mod(t + 0.5 * hs, period) < length, amp, 0.0)
        f0, aux = rhs(y, g, stim)
        colV, rowV, rowC, diag = jacobian(y, g, stim, f0, aux)
        a = 1.0 / (GAM * hs)
        g1 = solve(a, colV, rowV, rowC, diag, f0)
        f2, _ = rhs(y + A21 * g1, g, stim)
        g2 = solve(a, colV, rowV, rowC, diag, f2 + C21 * g1 / hs)
        f3, _ = rhs(y + A31 * g1 + A32 * g2, g, stim)
        g3 = solve(a, colV, rowV, rowC, diag, f3 + (C31 * g1 + C32 * g2) / hs)
        g4 = solve(a, colV, rowV, rowC, diag, f3 + (C41 * g1 + C42 * g2 + C43 * g3) / hs)
        yn = y + B1 * g1 + B2 * g2 + B3 * g3 + B4 * g4
        err = E1 * g1 + E2 * g2 + E3 * g3 + E4 * g4
        sc = atol * w + rtol * np.abs(y)
        e2 = np.mean((err / sc) ** 2, axis=0)
        ok = active & (e2 <= 1.0) & np.isfinite(e2)
        with np.errstate(divide='ignore', invalid='ignore'):
            fac = np.clip(0.9 * e2 ** (-1.0 / 8.0), 0.2, 5.0)
        fac = np.where(np.isfinite(fac), fac, 0.2)
        tn = np.where(clipped, edge, t + hs)                 # land exactly on the edge
        # dense output (cubic Hermite) needs f(yn): evaluated with the NEXT interval's stimulus
        stim_n = np.where(np.mod(tn + 1e-9, period) < length, amp, 0.0)
        fn_left, _ = rhs(yn, g, stim)                        # same interval (left limit)
        idx = np.where(ok)[0]
        for i in idx:
            k = nobs[i]
            while k < len(times) and times[k] <= tn[i]:
                th = (times[k] - t[i]) / hs[i]
                for o in range(2):
                    y0_, y1_ = y[o, i], yn[o, i]
                    d0, d1 = hs[i] * f0[o, i], hs[i] * fn_left[o, i]
                    h00 = (1 + 2 * th) * (1 - th) ** 2; h10 = th * (1 - th) ** 2
                    h01 = th * th * (3 - 2 * th); h11 = th * th * (th - 1)
                    out[i, k, o] = h00 * y0_ + h10 * d0 + h01 * y1_ + h11 * d1
                k += 1
            nobs[i] = k
        accepted += ok
        t = np.where(ok, tn, t)
        y = np.where(ok, yn, y)
        # a clipped, accepted step keeps the unclipped proposal
        h = np.where(ok & clipped, np.maximum(h, hs * fac), hs * fac)
        h = np.where(ok & clipped & (np.mod(t, period) < length), np.minimum(h, length), h)
    return out, attempts, accepted


def scores_of(traj, values, sigma_cols):
    # Gaussian log-likelihood with per-output sigma taken from the parameter vector
    n_t = values.shape[0]
    ll = np.zeros(traj.shape[0])
    for o in range(2):
        s = sigma_cols[:, o]
        ss = np.sum((values[None, :, o] - traj[:, :, o]) ** 2, axis=1)
        ll += -0.5 * n_t * np.log(2 * np.pi) - n_t * np.log(s) - ss / (2 * s * s)
    return ll


if __name__ == '__main__':
    gld = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'beeler_reuter.npz'))
    xs, times, truth = gld['xs'], gld['times'], gld['truth']
    ref_err = np.abs(gld['traj'] - truth)
    print('reference (LSODA 1e-4/1e-6): V err %.3g  Cai rel err %.3g' %
          (ref_err[..., 0].max(), (ref_err[..., 1] / truth[..., 1]).max()))
    for tol in (1e-4, 1e-5, 1e-6, 1e-7, 1e-8):
        traj, att, acc = integrate(xs[:8], times, tol, tol)
        e = np.abs(traj - truth)
        ll = scores_of(traj, gld['values'], xs[:8, 5:7])
        rel = np.abs(ll - gld['truth_scores'][:8]) / np.abs(gld['truth_scores'][:8])
        print('tol %.0e: attempts %d..%d (accepted %d)  V err %.3g mV  Cai rel %.3g  score rel %.3g' %
              (tol, att.min(), att.max(), acc.max(), e[..., 0].max(), (e[..., 1] / truth[..., 1]).max(), rel.max()))
