"""keras.regularizers: the reference only ever passes None on this path.  TEST INFRASTRUCTURE."""


def get(identifier):
    if identifier is None:
        return None
    raise NotImplementedError('regularizers are not used by the LSTUR path')


def serialize(r):
    return None
