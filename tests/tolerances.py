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


def elem_violations(a, b, tol=1e-3):
    """fraction of elements with |a-b| > tol*|b| + tol*rms(b)"""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    rms = float(np.sqrt(np.mean(b * b)))
    return float(np.mean(np.abs(a - b) > tol * np.abs(b) + tol * rms))


def assert_close(a, b, tol=1e-3, name='', tail=0.0, tail_excess=1.0):
    """both forms of the 1e-3 bound.  tail > 0 (large samples only): at most that fraction of the elements may exceed the
    element-wise bound, and none by more than the factor tail_excess — measured on B200 at C3 scale: with 16-bit operands
    1 of 640 000 history-vector elements exceeds it, by 5 %."""
    r, x = rel(a, b), elem_excess(a, b, tol)
    assert r < tol, '%s: norm-wise relative error %.3e >= %.1e' % (name, r, tol)
    if tail > 0:
        v = elem_violations(a, b, tol)
        assert v <= tail and x <= tail_excess, '%s: element-wise bound: %.2e of the elements violate it, worst by a factor %.2f' % (name, v, x)
    else:
        assert x <= 1.0, '%s: element-wise |a-b| <= tol*|b| + tol*rms(b) violated by a factor %.2f' % (name, x)


def assert_adam_weights_close(w, w_ref, tol, lr, steps, name='', exact=False, tail=1e-3):
    """Weights after `steps` Keras-Adam steps, started from the same point, from gradients that agree to rounding.

    Adam normalises every element's gradient to ~lr, so an element whose gradient is at the level of the rounding noise
    (a dead ReLU feature, a word row seen once) takes a +-lr step whose SIGN the noise decides: a 1e-7 difference in the
    gradients becomes up to 2*lr per step in that one weight.  `exact` (fp32 verification mode): every element within
    `tol`.  Otherwise: at most `tail` of the elements may exceed `tol`, none may differ by more than the 2*lr*steps two
    opposite runs of Adam steps can produce — a wrong exchange (a missing rank, a wrong scale) moves EVERY element by a
    fraction of lr and fails the first form."""
    d = np.abs(np.asarray(w, dtype=np.float64) - np.asarray(w_ref, dtype=np.float64))
    worst = float(d.max()) if d.size else 0.0
    if exact:
        assert worst <= tol, (name, worst)
        return worst, 0.0
    frac = float(np.mean(d > tol)) if d.size else 0.0
    assert frac <= tail and worst <= 2.02 * lr * steps, '%s: %.2e of the elements differ by more than %.1e, worst %.3e' % (name, frac, tol, worst)
    return worst, frac
