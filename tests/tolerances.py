"""Tolerance forms used by the GPU parity tests (BASELINE.json north_star: "within 1e-3 relative on encoder outputs,
scores and loss").

rel()          norm-wise:     max|a-b| / max|b|
elem_excess()  element-wise:  max over elements of |a-b| / (tol*|b| + tol*rms(b)); <= 1 passes.  The rms floor keeps
               elements that are ~0 by cancellation from demanding absolute accuracy below the tensor's own scale.
"""
import numpy as np


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def elem_excess(a, b, tol=1e-3):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    rms = float(np.sqrt(np.mean(b * b)))
    return float((np.abs(a - b) / (tol * np.abs(b) + tol * rms + 1e-300)).max())


def assert_close(a, b, tol=1e-3, name=''):
    """both forms of the 1e-3 bound"""
    r, x = rel(a, b), elem_excess(a, b, tol)
    assert r < tol, '%s: norm-wise relative error %.3e >= %.1e' % (name, r, tol)
    assert x <= 1.0, '%s: element-wise |a-b| <= tol*|b| + tol*rms(b) violated by a factor %.2f' % (name, x)
