"""Restatement of the part of stand-alone Keras 2.2.x (TensorFlow-1.x backend) that the LSTUR path of nvagus/mnexp calls,
on torch.  TEST INFRASTRUCTURE ONLY — see oracle/keras_shim/README.md."""
from . import activations, backend, callbacks, constraints, initializers, layers, losses, metrics, models, optimizers, regularizers, utils  # noqa: F401
from ._engine import Input, Model                  # noqa: F401
from .models import Sequential                     # noqa: F401

__version__ = '2.2.4'
