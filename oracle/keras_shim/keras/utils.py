"""keras.utils.  TEST INFRASTRUCTURE."""
import numpy as np


def to_categorical(y, num_classes=None, dtype='float32'):
    y = np.array(y, dtype='int')
    shape = y.shape
    if shape and shape[-1] == 1 and len(shape) > 1:
        shape = tuple(shape[:-1])
    y = y.ravel()
    if not num_classes:
        num_classes = np.max(y) + 1
    out = np.zeros((y.shape[0], num_classes), dtype=dtype)
    out[np.arange(y.shape[0]), y] = 1
    return out.reshape(shape + (num_classes,))
