"""keras.activations.  TEST INFRASTRUCTURE (oracle/keras_shim/README.md)."""
from . import backend as K


def linear(x):
    return x


def relu(x, alpha=0., max_value=None):
    return K.relu(x, alpha, max_value)


def tanh(x):
    return K.tanh(x)


def sigmoid(x):
    return K.sigmoid(x)


def hard_sigmoid(x):
    return K.hard_sigmoid(x)


def softmax(x, axis=-1):
    return K.softmax(x, axis)


def get(identifier):
    if identifier is None:
        return linear
    if callable(identifier):
        return identifier
    return globals()[identifier]


def serialize(fn):
    return fn.__name__
