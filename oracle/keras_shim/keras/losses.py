"""keras.losses.  TEST INFRASTRUCTURE (oracle/keras_shim/README.md)."""
from . import backend as K


def categorical_crossentropy(y_true, y_pred):
    return K.categorical_crossentropy(y_true, y_pred)


def binary_crossentropy(y_true, y_pred):
    return K.mean(K.binary_crossentropy(y_true, y_pred), axis=-1)


def mean_squared_error(y_true, y_pred):
    return K.mean(K.square(K._t(y_pred) - K._t(y_true)), axis=-1)


mse = MSE = mean_squared_error


def get(identifier):
    if identifier is None or callable(identifier):
        return identifier
    return globals()[identifier]
