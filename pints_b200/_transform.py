"""
Search space -> model space for a whole batch of positions.

The reference's controllers wrap the objective when they are given a
``transformation`` (pints/_optimisers/__init__.py:392-406,
pints/_mcmc/__init__.py:349-361): ``TransformedErrorMeasure``,
``TransformedLogLikelihood`` (``f(to_model(q))``, pints/_transformation.py:
1060-1064, :1280-1282) or ``TransformedLogPDF`` (``f(to_model(q)) +
log_jacobian_det(q)``, :1159-1169).  The forward solve wants model-space
parameters, so ``CudaEvaluator`` unwraps the transformation, maps the rows
here on the host -- O(population * d), nothing to solve -- and launches the
device path on the result; the Jacobian term is added to the scores
afterwards.

For the reference's element-wise transformations the rows are mapped with the
same NumPy expressions the reference applies to one vector, on the whole
(n, d) array (identical results element by element); any other transformation
object is applied row by row through its own ``to_model`` /
``log_jacobian_det``, i.e. exactly the reference's arithmetic.
"""
import numpy as np


def _cls(t):
    mod = type(t).__module__ or ''
    return type(t).__name__ if mod.split('.')[0] == 'pints' else None


def _expit(q):
    from scipy.special import expit
    return expit(q)


def _vectorised_to_model(t, q):
    """(n, d) -> (n, d) for the reference's element-wise classes, or None."""
    kind = _cls(t)
    if kind == 'IdentityTransformation':
        return q
    if kind == 'LogTransformation':                      # :722-725
        return np.exp(q)
    if kind == 'LogitTransformation':                    # :639-642
        return _expit(q)
    if kind == 'RectangularBoundariesTransformation':    # :846-849
        return (t._b - t._a) * _expit(q) + t._a
    if kind in ('ScalingTransformation', 'UnitCubeTransformation'):   # :914-919
        p = t._inv_s * q
        if t._translation is not None:
            p -= t._translation
        return p
    if kind == 'ComposedTransformation':                 # :402-411
        out = np.zeros(q.shape)
        hi = 0
        for sub in t._transformations:
            lo, hi = hi, hi + sub.n_parameters()
            part = _vectorised_to_model(sub, q[:, lo:hi])
            if part is None:
                return None
            out[:, lo:hi] = part
        return out
    return None


def _row_sum(a):
    """Row sums in the order ``np.sum`` adds a short 1-d vector (sequential for
    fewer than 8 entries), so that they equal the reference's per-vector sums
    bit for bit; None for longer rows (pairwise blocks there)."""
    n, d = a.shape
    if d == 0:
        return np.zeros(n)
    if d >= 8:
        return None
    acc = a[:, 0].copy()
    for k in range(1, d):
        acc = acc + a[:, k]
    return acc


def _vectorised_log_jacobian_det(t, q):
    """(n, d) -> (n,) for the reference's element-wise classes, or None."""
    kind = _cls(t)
    n = q.shape[0]
    if kind == 'IdentityTransformation':                 # :543-545
        return np.zeros(n)
    if kind == 'LogTransformation':                      # :706-709
        return _row_sum(q)
    if kind == 'LogitTransformation':                    # :623-626
        e = _expit(q)
        return _row_sum(np.log(e) + np.log(1. - e))
    if kind == 'RectangularBoundariesTransformation':    # :829-833, :856-858
        s = np.log(1. + np.exp(-q))
        return _row_sum(np.log(t._b - t._a) - 2. * s - q)
    if kind in ('ScalingTransformation', 'UnitCubeTransformation'):   # :902-904
        return np.full(n, np.sum(np.log(np.abs(t._inv_s))))
    if kind == 'ComposedTransformation' and getattr(t, '_elementwise', False):
        # :441-453 (element-wise case): 0 + term + term + ..., in order
        total = np.zeros(n)
        hi = 0
        for sub in t._transformations:
            lo, hi = hi, hi + sub.n_parameters()
            part = _vectorised_log_jacobian_det(sub, q[:, lo:hi])
            if part is None:
                return None
            total = total + part
        return total
    return None


def rows_to_model(transformation, q):
    """``transformation.to_model`` of every row of ``q`` (n, d)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    p = _vectorised_to_model(transformation, q)
    if p is None:
        p = np.array([np.asarray(transformation.to_model(row), dtype=float)
                      for row in q]).reshape(q.shape)
    return np.ascontiguousarray(p, dtype=np.float64)


def rows_log_jacobian_det(transformation, q):
    """``transformation.log_jacobian_det`` of every row of ``q`` (n, d)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = _vectorised_log_jacobian_det(transformation, q)
    if out is None:
        out = np.array([float(transformation.log_jacobian_det(row))
                        for row in q])
    return out
