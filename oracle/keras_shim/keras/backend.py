"""keras.backend (TensorFlow-1.x backend semantics) on torch — the functions the reference calls.  TEST INFRASTRUCTURE."""
import numpy as np
import torch

from . import _engine as E
from ._engine import floatx, unwrap as _v, variable


def backend():
    return 'tensorflow'


def set_floatx(name):
    E._STATE['floatx'] = str(name)


def epsilon():
    return E._STATE['epsilon']


def set_epsilon(e):
    E._STATE['epsilon'] = float(e)


def learning_phase():
    return E._STATE['phase']


def set_learning_phase(v):
    E._STATE['phase'] = int(v)


def in_train_phase(x, alt, training=None):
    training = learning_phase() if training is None else training
    pick = x if training else alt
    return pick() if callable(pick) else pick


def _t(x, like=None):
    x = _v(x)
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(x, dtype=like.dtype if isinstance(like, torch.Tensor) and like.is_floating_point() else None)


def dtype(x):
    return str(_t(x).dtype).replace('torch.', '')


def cast(x, dt):
    return _t(x).to(E.tdtype(dt))


def get_value(x):
    return _t(x).detach().cpu().numpy().copy()


def set_value(x, value):
    x.data.copy_(torch.as_tensor(np.asarray(value), dtype=x.dtype))


def batch_get_value(xs):
    return [get_value(x) for x in xs]


def batch_set_value(pairs):
    for x, v in pairs:
        set_value(x, v)


def eval(x):
    return get_value(x)


def constant(value, dtype=None, shape=None):
    t = torch.as_tensor(np.asarray(value), dtype=E.tdtype(dtype))
    return t.expand(shape).clone() if shape is not None else t


def zeros_like(x, dtype=None, name=None):
    x = _t(x)
    return torch.zeros_like(x, dtype=E.tdtype(dtype) if dtype else x.dtype)


def ones_like(x, dtype=None, name=None):
    x = _t(x)
    return torch.ones_like(x, dtype=E.tdtype(dtype) if dtype else x.dtype)


def zeros(shape, dtype=None, name=None):
    return variable(np.zeros(shape), dtype, name)


def ones(shape, dtype=None, name=None):
    return variable(np.ones(shape), dtype, name)


def shape(x):
    return tuple(_t(x).shape)


def int_shape(x):
    return x._keras_shape if isinstance(x, E.Sym) else tuple(_t(x).shape)


def ndim(x):
    return _t(x).dim()


def expand_dims(x, axis=-1):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis):
    return _t(x).squeeze(axis)


def reshape(x, shape):
    return _t(x).reshape(tuple(int(s) for s in shape))


def permute_dimensions(x, pattern):
    return _t(x).permute(*pattern)


def concatenate(tensors, axis=-1):
    return torch.cat([_t(t) for t in tensors], dim=axis)


def stack(xs, axis=0):
    return torch.stack([_t(t) for t in xs], dim=axis)


def tile(x, n):
    return _t(x).repeat(*n) if isinstance(n, (list, tuple)) else _t(x).repeat(n)


def repeat_elements(x, rep, axis):
    return torch.repeat_interleave(_t(x), rep, dim=axis)


def _axis(axis):
    return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axis(axis), keepdim=keepdims)


def mean(x, axis=None, keepdims=False):
    x = _t(x)
    if not x.is_floating_point():
        x = x.to(E.tdtype())
    return x.mean() if axis is None else x.mean(dim=_axis(axis), keepdim=keepdims)


def max(x, axis=None, keepdims=False):
    x = _t(x)
    return x.max() if axis is None else x.amax(dim=_axis(axis), keepdim=keepdims)


def min(x, axis=None, keepdims=False):
    x = _t(x)
    return x.min() if axis is None else x.amin(dim=_axis(axis), keepdim=keepdims)


def any(x, axis=None, keepdims=False):
    x = _t(x).bool()
    return x.any() if axis is None else x.any(dim=axis, keepdim=keepdims)


def all(x, axis=None, keepdims=False):
    x = _t(x).bool()
    return x.all() if axis is None else x.all(dim=axis, keepdim=keepdims)


def argmax(x, axis=-1):
    return _t(x).argmax(dim=axis)


def equal(x, y):
    x = _t(x)
    return x == _t(y, x)


def not_equal(x, y):
    x = _t(x)
    return x != _t(y, x)


def greater(x, y):
    x = _t(x)
    return x > _t(y, x)


def greater_equal(x, y):
    x = _t(x)
    return x >= _t(y, x)


def less(x, y):
    x = _t(x)
    return x < _t(y, x)


def maximum(x, y):
    x = _t(x)
    return torch.maximum(x, _t(y, x).to(x.dtype))


def minimum(x, y):
    x = _t(x)
    return torch.minimum(x, _t(y, x).to(x.dtype))


def switch(condition, then_expression, else_expression):
    c = _t(condition)
    a = then_expression() if callable(then_expression) else _t(then_expression)
    b = else_expression() if callable(else_expression) else _t(else_expression)
    while c.dim() < a.dim():
        c = c.unsqueeze(-1)
    return torch.where(c.bool(), a, b)


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def sqrt(x):
    return torch.sqrt(torch.clamp(_t(x), min=0.))


def square(x):
    return _t(x) ** 2


def abs(x):
    return torch.abs(_t(x))


def pow(x, a):
    return _t(x) ** a


def clip(x, lo, hi):
    return torch.clamp(_t(x), lo, hi)


def round(x):
    return torch.round(_t(x))


def tanh(x):
    return torch.tanh(_t(x))


def sigmoid(x):
    return torch.sigmoid(_t(x))


def hard_sigmoid(x):
    return torch.clamp(0.2 * _t(x) + 0.5, 0., 1.)


def relu(x, alpha=0., max_value=None):
    x = _t(x)
    y = torch.where(x > 0, x, alpha * x) if alpha else torch.relu(x)
    return torch.clamp(y, max=max_value) if max_value is not None else y


def softmax(x, axis=-1):
    return torch.softmax(_t(x), dim=axis)


def dot(x, y):
    return torch.matmul(_t(x), _t(y))


def bias_add(x, bias, data_format=None):
    return _t(x) + _t(bias)


def batch_dot(x, y, axes=None):
    """tensorflow_backend.batch_dot for the ranks the reference uses (2-D and 3-D operands)"""
    x, y = _t(x), _t(y)
    if isinstance(axes, int):
        axes = (axes, axes)
    if axes is None:
        axes = (x.dim() - 1, y.dim() - 2)
    a0 = axes[0] % x.dim()
    a1 = axes[1] % y.dim()
    if x.dim() == 2 and y.dim() == 2:
        if a0 != 1 or a1 != 1:
            raise ValueError('batch_dot: cannot reduce the batch axis')
        return (x * y).sum(1, keepdim=True)
    xd, yd = x.dim(), y.dim()
    if xd < yd:
        x = x.reshape(tuple(x.shape) + (1,) * (yd - xd))
    elif yd < xd:
        y = y.reshape(tuple(y.shape) + (1,) * (xd - yd))
    # move the reduced axis: last of x, second of y (tf.matmul with adjoint flags)
    xm = x.movedim(a0, -1) if x.dim() == 3 else x
    ym = y.movedim(a1, 1) if y.dim() == 3 else y
    out = torch.matmul(xm, ym)
    if xd != yd:
        idx = (xd + yd - 3) if xd > yd else (xd - 1)
        out = out.squeeze(idx)
    if out.dim() == 1:
        out = out.unsqueeze(1)
    return out


def gather(reference, indices):
    return _t(reference)[_t(indices).long()]


def categorical_crossentropy(target, output, from_logits=False, axis=-1):
    target, output = _t(target), _t(output)
    if from_logits:
        return -(target * torch.log_softmax(output, dim=axis)).sum(axis)
    output = output / output.sum(axis, keepdim=True)
    output = torch.clamp(output, epsilon(), 1. - epsilon())
    return -(target * torch.log(output)).sum(axis)


def binary_crossentropy(target, output, from_logits=False):
    target, output = _t(target), _t(output)
    if not from_logits:
        output = torch.clamp(output, epsilon(), 1. - epsilon())
        output = torch.log(output / (1. - output))
    return torch.nn.functional.binary_cross_entropy_with_logits(output, target, reduction='none')


def dropout(x, level, noise_shape=None, seed=None):
    """tf.nn.dropout: keep with probability 1-level, scale kept values by 1/(1-level).  TensorFlow's random stream cannot
    be reproduced; the fixture generator installs a hook that supplies the keep mask (oracle/keras_shim/README.md)."""
    x = _t(x)
    shp = list(x.shape)
    if noise_shape is not None:
        shp = [s if n is None else n for s, n in zip(shp, noise_shape)]
    hook = E._STATE['dropout_hook']
    keep = hook(tuple(shp), level) if hook else (torch.rand(shp) >= level)
    keep = torch.as_tensor(keep).to(x.dtype)
    return x * keep / (1. - level)


def get_session():
    class _Session:
        def run(self, *a, **k):
            return None
    return _Session()


def clear_session():
    E.reset_uids()


def stop_gradient(x):
    return _t(x).detach()


def is_keras_tensor(x):
    return isinstance(x, E.Sym)
