"""keras.initializers (numpy draws; fixtures overwrite the weights anyway).  TEST INFRASTRUCTURE."""
import numpy as np

_rng = np.random.RandomState(20190131)


def _fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    if len(shape) in (3, 4, 5):           # conv kernels: (spatial..., in, out)
        rf = int(np.prod(shape[:-2]))
        return shape[-2] * rf, shape[-1] * rf
    n = int(np.sqrt(np.prod(shape))) if shape else 1
    return n, n


class Initializer:
    def get_config(self):
        return {}


class Zeros(Initializer):
    def __call__(self, shape, dtype=None):
        return np.zeros(shape)


class Ones(Initializer):
    def __call__(self, shape, dtype=None):
        return np.ones(shape)


class Constant(Initializer):
    def __init__(self, value=0):
        self.value = value

    def __call__(self, shape, dtype=None):
        return np.full(shape, self.value, dtype=np.float64)


class RandomUniform(Initializer):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval = minval, maxval

    def __call__(self, shape, dtype=None):
        return _rng.uniform(self.minval, self.maxval, shape)


class RandomNormal(Initializer):
    def __init__(self, mean=0., stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape, dtype=None):
        return _rng.normal(self.mean, self.stddev, shape)


class VarianceScaling(Initializer):
    def __init__(self, scale=1.0, mode='fan_in', distribution='normal', seed=None):
        self.scale, self.mode, self.distribution = scale, mode, distribution

    def __call__(self, shape, dtype=None):
        fi, fo = _fans(shape)
        s = self.scale / max(1., {'fan_in': fi, 'fan_out': fo, 'fan_avg': (fi + fo) / 2.}[self.mode])
        if self.distribution == 'normal':
            return _rng.normal(0., np.sqrt(s) / .87962566103423978, shape).clip(-2 * np.sqrt(s), 2 * np.sqrt(s))
        lim = np.sqrt(3. * s)
        return _rng.uniform(-lim, lim, shape)


class Orthogonal(Initializer):
    def __init__(self, gain=1., seed=None):
        self.gain = gain

    def __call__(self, shape, dtype=None):
        rows = int(np.prod(shape[:-1]))
        a = _rng.normal(0., 1., (rows, shape[-1]))
        u, _, v = np.linalg.svd(a, full_matrices=False)
        q = u if u.shape == (rows, shape[-1]) else v
        return self.gain * q.reshape(shape)


def glorot_uniform(seed=None):
    return VarianceScaling(1., 'fan_avg', 'uniform')


def glorot_normal(seed=None):
    return VarianceScaling(1., 'fan_avg', 'normal')


def he_normal(seed=None):
    return VarianceScaling(2., 'fan_in', 'normal')


zeros, ones, constant, uniform, normal, orthogonal = Zeros, Ones, Constant, RandomUniform, RandomNormal, Orthogonal
random_uniform, random_normal = RandomUniform, RandomNormal


def get(identifier):
    if identifier is None:
        return None
    if isinstance(identifier, str):
        obj = globals()[identifier]
        return obj() if isinstance(obj, type) or identifier.startswith(('glorot', 'he_')) else obj
    if isinstance(identifier, type):
        return identifier()
    return identifier


def serialize(i):
    return i.__class__.__name__
