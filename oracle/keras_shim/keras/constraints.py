"""keras.constraints: applied to a weight after every optimizer update.  TEST INFRASTRUCTURE."""
import torch

from . import backend as K


class Constraint:
    def __call__(self, w):
        return w


class NonNeg(Constraint):
    def __call__(self, w):
        return w * (w >= 0.).to(w.dtype)


class MinMaxNorm(Constraint):
    def __init__(self, min_value=0.0, max_value=1.0, rate=1.0, axis=0):
        self.min_value, self.max_value, self.rate, self.axis = min_value, max_value, rate, axis

    def __call__(self, w):
        norms = torch.sqrt((w ** 2).sum(dim=self.axis, keepdim=True))
        desired = self.rate * torch.clamp(norms, self.min_value, self.max_value) + (1 - self.rate) * norms
        return w * (desired / (K.epsilon() + norms))


non_neg, min_max_norm = NonNeg, MinMaxNorm


def get(identifier):
    if identifier is None or isinstance(identifier, Constraint) or callable(identifier):
        return identifier
    return globals()[identifier]()


def serialize(c):
    return None if c is None else c.__class__.__name__
