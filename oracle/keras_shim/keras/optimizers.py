"""keras.optimizers (2.2.4).  TEST INFRASTRUCTURE (oracle/keras_shim/README.md).

Embedding gradients arrive as tf.IndexedSlices in the real thing and `(1 - beta) * g` densifies them, so Adam moves
EVERY row of an embedding table every step (m, v decay) — the dense update below."""
import numpy as np
import torch

from . import backend as K


class Optimizer:
    def __init__(self, **kwargs):
        self.clipnorm, self.clipvalue = kwargs.get('clipnorm'), kwargs.get('clipvalue')

    def _constrain(self, p):
        if getattr(p, 'constraint', None) is not None:
            p.data.copy_(p.constraint(p.data))


class SGD(Optimizer):
    def __init__(self, lr=0.01, momentum=0., decay=0., nesterov=False, **kwargs):
        super().__init__(**kwargs)
        self.lr, self.momentum, self.decay, self.nesterov = K.variable(lr), momentum, decay, nesterov
        self.iterations, self.moments = 0, {}

    def apply(self, params, grads):
        lr = float(self.lr.detach()) * (1. / (1. + self.decay * self.iterations)) if self.decay > 0 else float(self.lr.detach())
        self.iterations += 1
        with torch.no_grad():
            for p, g in zip(params, grads):
                m = self.moments.setdefault(id(p), torch.zeros_like(p))
                v = self.momentum * m - lr * g
                m.copy_(v)
                p.add_(self.momentum * v - lr * g if self.nesterov else v)
                self._constrain(p)


class Adam(Optimizer):
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0., amsgrad=False, **kwargs):
        super().__init__(**kwargs)
        self.lr, self.beta_1, self.beta_2, self.decay = K.variable(lr, dtype='float64'), beta_1, beta_2, decay
        self.epsilon = K.epsilon() if epsilon is None else epsilon
        self.initial_decay, self.amsgrad = decay, amsgrad
        self.iterations, self.ms, self.vs = 0, {}, {}

    def apply(self, params, grads):
        lr = float(self.lr.detach())
        if self.initial_decay > 0:
            lr = lr * (1. / (1. + self.decay * self.iterations))
        t = self.iterations + 1
        lr_t = lr * (np.sqrt(1. - self.beta_2 ** t) / (1. - self.beta_1 ** t))
        with torch.no_grad():
            for p, g in zip(params, grads):
                m = self.ms.setdefault(id(p), torch.zeros_like(p))
                v = self.vs.setdefault(id(p), torch.zeros_like(p))
                m.mul_(self.beta_1).add_((1. - self.beta_1) * g)
                v.mul_(self.beta_2).add_((1. - self.beta_2) * g * g)
                p.sub_(lr_t * m / (torch.sqrt(v) + self.epsilon))
                self._constrain(p)
        self.iterations += 1


sgd, adam = SGD, Adam


def get(identifier):
    if isinstance(identifier, Optimizer):
        return identifier
    return {'sgd': SGD, 'adam': Adam}[identifier.lower()]()
