"""keras.metrics.  TEST INFRASTRUCTURE (oracle/keras_shim/README.md)."""
from . import backend as K
from . import losses
from .losses import binary_crossentropy, categorical_crossentropy, mean_squared_error     # noqa: F401


def categorical_accuracy(y_true, y_pred):
    return K.cast(K.equal(K.argmax(y_true, axis=-1), K.argmax(y_pred, axis=-1)), K.floatx())


def binary_accuracy(y_true, y_pred):
    return K.mean(K.cast(K.equal(K._t(y_true), K.round(y_pred)), K.floatx()), axis=-1)


def get(identifier, loss_fn=None):
    if callable(identifier):
        return identifier
    if identifier in ('accuracy', 'acc'):
        if loss_fn is losses.categorical_crossentropy:
            return categorical_accuracy
        return binary_accuracy
    return globals()[identifier]
