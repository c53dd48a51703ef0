"""The handful of TensorFlow-1.x functions the LSTUR path of nvagus/mnexp calls directly (tf.log, tf.reduce_mean,
tf.gfile, ...), on torch / the local file system.  TEST INFRASTRUCTURE ONLY — see oracle/keras_shim/README.md."""
import builtins
import os

import torch

from keras import backend as _K
from keras._engine import unwrap as _v

float32, float64, int32, int64 = 'float32', 'float64', 'int32', 'int64'
__version__ = '1.12.0'


def log(x):
    return torch.log(_v(x))


def exp(x):
    return torch.exp(_v(x))


def tanh(x):
    return torch.tanh(_v(x))


def round(x):
    return torch.round(_v(x))


def identity(x):
    return x


def reduce_mean(x, axis=None, keepdims=False):
    return _K.mean(x, axis, keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    return _K.sum(x, axis, keepdims)


def reshape(x, shape):
    return _v(x).reshape(tuple(int(s) for s in shape))


def unstack(x, num=None, axis=0):
    return list(torch.unbind(_v(x), dim=axis))


def constant(value, dtype=None, shape=None):
    return _K.constant(value, dtype, shape)


def Variable(initial_value, constraint=None, **kw):
    return _K.variable(initial_value, constraint=constraint)


def local_variables():
    return []


def add_to_collection(name, value):
    pass


class GraphKeys:
    GLOBAL_VARIABLES = 'variables'


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def control_dependencies(ops):
    return _Null()


class initializers:
    @staticmethod
    def variables(var_list):
        return None


class errors:
    NotFoundError = FileNotFoundError


class gfile:
    """tf.gfile on the local file system"""

    @staticmethod
    def GFile(name, mode='r'):
        return builtins.open(name, mode)

    Open = GFile

    @staticmethod
    def MkDir(path):
        os.makedirs(path, exist_ok=True)

    MakeDirs = MkDir

    @staticmethod
    def ListDirectory(path):
        return os.listdir(path)

    @staticmethod
    def Exists(path):
        return os.path.exists(path)


class metrics:
    @staticmethod
    def auc(labels, predictions, **kw):
        """tf.metrics.auc is a STREAMING metric (200-bin Riemann sum over all batches so far); only used as a progress
        display by the sigmoid-family tasks — here the exact AUC of the current batch"""
        y, p = _v(labels).reshape(-1).double(), _v(predictions).reshape(-1).double()
        pos, neg = p[y > 0.5], p[y <= 0.5]
        if len(pos) == 0 or len(neg) == 0:
            value = torch.tensor(0.)
        else:
            d = pos[:, None] - neg[None, :]
            value = ((d > 0).double() + 0.5 * (d == 0).double()).mean()
        return value, None
