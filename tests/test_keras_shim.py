"""Known-answer tests of oracle/keras_shim — the restatement of Keras 2.2.x under which the reference's own files are run
(tests/golden/make_ref_golden.py).  Every expectation below is a documented behaviour of stand-alone Keras 2.2.4 with the
TensorFlow backend (file named per test), computed by hand here; none of them is taken from this repo's oracle, so the two
restatements (shim layers, oracle closed forms) stay independent.

The shim is imported under its path name (oracle.keras_shim.keras), not as `keras`, so nothing is shadowed in the test
process."""
import numpy as np
import pytest
import torch

from oracle.keras_shim import keras
from oracle.keras_shim.keras import backend as K
from oracle.keras_shim.keras import layers as L


@pytest.fixture(autouse=True)
def _float64():
    K.set_floatx('float64')
    K.clear_session()
    yield
    K.set_floatx('float32')


def t(a):
    return torch.tensor(np.asarray(a, dtype=np.float64))


def test_masking_flags_and_zeroes_all_zero_steps():
    """layers/core.py Masking: mask = any(x != 0, -1); output = x * mask"""
    x = keras.Input((3, 2))
    m = keras.Model(x, L.Masking()(x))
    v = np.array([[[0, 0], [1, 0], [0, 0]]], dtype=np.float64)
    assert np.array_equal(m.predict(v), v)
    sym = L.Masking()(x)
    assert sym.mask_example is not None and tuple(sym.mask_example.shape) == (2, 3)


def test_layer_without_mask_support_rejects_a_mask():
    """engine/base_layer.py Layer.compute_mask: TypeError when a mask reaches a layer that does not support masking"""
    x = keras.Input((3, 2))
    masked = L.Masking()(x)
    with pytest.raises(TypeError, match='does not support masking'):
        L.Reshape((6,))(masked)


def test_lambda_drops_the_mask_and_dense_dropout_pass_it():
    x = keras.Input((3, 2))
    masked = L.Masking()(x)
    assert L.Lambda(lambda z: z * 2)(masked).mask_example is None          # layers/core.py Lambda.compute_mask -> self.mask (None)
    assert L.Dense(4)(masked).mask_example is not None                     # Dense.supports_masking = True
    assert L.Dropout(0.5)(masked).mask_example is not None


def test_gru_cell_and_masked_steps_by_hand():
    """layers/recurrent.py GRUCell (implementation 1, reset_after=False) + backend.rnn masking: a masked step keeps the
    state; the layer returns the last state; an initial_state is the state before step 0"""
    u = 1
    x = keras.Input((3, 1))
    h0 = keras.Input((1,))
    gru = L.GRU(u)
    out = gru(L.Masking()(x), initial_state=h0)
    m = keras.Model([x, h0], out)
    gru.set_weights([np.array([[0.5, -0.3, 0.8]]), np.array([[0.2, 0.4, -0.6]]), np.array([0.1, 0.0, -0.1])])
    hs = lambda v: min(1.0, max(0.0, 0.2 * v + 0.5))

    def step(xv, h):
        z = hs(0.5 * xv + 0.1 + 0.2 * h)
        r = hs(-0.3 * xv + 0.0 + 0.4 * h)
        hh = np.tanh(0.8 * xv - 0.1 + (-0.6) * (r * h))
        return z * h + (1 - z) * hh
    seq, start = [1.5, 0.0, -2.0], 0.3                 # the middle step is all-zero -> masked
    h = step(seq[0], start)
    h = step(seq[2], h)
    got = m.predict([np.array(seq).reshape(1, 3, 1), np.array([[start]])])
    assert abs(got[0, 0] - h) < 1e-12
    # every step masked: the initial state comes back
    got = m.predict([np.zeros((1, 3, 1)), np.array([[start]])])
    assert abs(got[0, 0] - start) < 1e-15


def test_conv1d_same_padding_and_kernel_layout():
    """layers/convolutional.py Conv1D: kernel (k, in, out); 'same' with k = 3 pads one step on each side"""
    x = keras.Input((4, 1))
    conv = L.Conv1D(1, 3, padding='same', activation='relu')
    m = keras.Model(x, conv(x))
    conv.set_weights([np.array([1.0, 10.0, 100.0]).reshape(3, 1, 1), np.array([0.5])])
    got = m.predict(np.array([1.0, 2.0, 3.0, 4.0]).reshape(1, 4, 1)).reshape(-1)
    assert np.allclose(got, [0 + 10 + 200 + 0.5, 1 + 20 + 300 + 0.5, 2 + 30 + 400 + 0.5, 3 + 40 + 0 + 0.5])


def test_time_distributed_model_and_nested_weights():
    inner_in = keras.Input((2,))
    dense = L.Dense(1, use_bias=False)
    inner = keras.Model(inner_in, dense(inner_in), name='inner')
    x = keras.Input((3, 2))
    m = keras.Model(x, L.TimeDistributed(inner)(x))
    dense.set_weights([np.array([[2.0], [-1.0]])])
    v = np.arange(6, dtype=np.float64).reshape(1, 3, 2)
    assert np.allclose(m.predict(v).reshape(-1), [2 * 0 - 1, 2 * 2 - 3, 2 * 4 - 5])
    assert len(m.trainable_weights) == 1 and m.trainable_weights[0] is dense.kernel
    inner.trainable = False
    assert m.trainable_weights == [] and len(m.non_trainable_weights) == 1


def test_dot_layer_shapes_and_dimension_check():
    """layers/merge.py Dot + backend.batch_dot: (B,T,D).(B,T) over axes (1,1) -> (B,D); 2-D operands -> (B,1); Dot.build
    rejects mismatched widths"""
    a, w = keras.Input((3, 2)), keras.Input((3,))
    m = keras.Model([a, w], L.Dot((1, 1))([a, w]))
    av = np.arange(6, dtype=np.float64).reshape(1, 3, 2)
    wv = np.array([[1.0, 10.0, 100.0]])
    assert np.allclose(m.predict([av, wv]), [[0 + 20 + 400, 1 + 30 + 500]])
    u, d = keras.Input((2,)), keras.Input((2,))
    assert keras.Model([u, d], L.dot([u, d], -1)).predict([np.array([[1.0, 2.0]]), np.array([[3.0, 4.0]])]).shape == (1, 1)
    with pytest.raises(ValueError, match='Dimension incompatibility'):
        L.dot([keras.Input((1,)), keras.Input((2,))], -1)


def test_model_layers_follow_keras_depth_order():
    """engine/network.py Network._init_graph_network: layers sorted by decreasing depth, ties by first visit in the
    depth-first walk from the outputs — the order of model.layers / get_weights() / the pkl files of utils.save_model"""
    a = keras.Input((2,), name='a')
    b = keras.Input((2,), name='b')
    da = L.Dense(2, name='da')(a)
    db = L.Dense(2, name='db')(b)
    cat = L.concatenate([db, da], name='cat')
    out = L.Dense(1, name='out')(cat)
    m = keras.Model([a, b], out)
    # depth 3: the inputs (b is reached first: concatenate lists db first), depth 2: db, da, depth 1: cat, depth 0: out
    assert [l.name for l in m.layers] == ['b', 'a', 'db', 'da', 'cat', 'out']
    shapes = [w.shape for w in m.get_weights()]
    assert shapes == [(2, 2), (2,), (2, 2), (2,), (4, 1), (1,)]


def test_categorical_crossentropy_renormalises_and_clips():
    """backend.categorical_crossentropy on probabilities: output /= sum; clip to [1e-7, 1 - 1e-7]; -sum(target * log)"""
    y = t([[1.0, 0.0]])
    assert abs(float(K.categorical_crossentropy(y, t([[2.0, 2.0]]))) - np.log(2.0)) < 1e-15
    assert abs(float(K.categorical_crossentropy(y, t([[0.0, 1.0]]))) + np.log(1e-7)) < 1e-12


def test_adam_first_steps_closed_form_and_dense_embedding_update():
    """optimizers.py Adam: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); p -= lr_t m / (sqrt(v) + 1e-7).  First step:
    -lr g / (|g| + 1e-7 / sqrt(1 - b2)) ... and an embedding row that got a gradient once keeps moving (dense update)"""
    x = keras.Input((1,))
    emb = L.Embedding(3, 1)
    m = keras.Model(x, L.Reshape((1,))(emb(x)))
    emb.set_weights([np.array([[1.0], [2.0], [3.0]])])
    m.compile(keras.optimizers.Adam(0.01), loss=lambda yt, yp: K.sum(yp, axis=-1))
    m.train_on_batch(np.array([1]), np.array([[0.0]]))          # d loss / d row1 = 1
    w = emb.get_weights()[0].reshape(-1)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    step1 = lr_t * 0.1 / (np.sqrt(0.001) + 1e-7)
    assert abs(w[1] - (2.0 - step1)) < 1e-15 and w[0] == 1.0 and w[2] == 3.0
    m.train_on_batch(np.array([2]), np.array([[0.0]]))          # row 1 gets no gradient now, its m / v decay but it still moves
    w2 = emb.get_weights()[0].reshape(-1)
    lr_t2 = 0.01 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    m1, v1 = 0.9 * 0.1, 0.999 * 0.001
    assert abs(w2[1] - (w[1] - lr_t2 * m1 / (np.sqrt(v1) + 1e-7))) < 1e-15 and w2[2] < 3.0


def test_dropout_only_in_training_and_hook_supplies_the_mask():
    from oracle.keras_shim.keras import _engine
    x = keras.Input((4,))
    m = keras.Model(x, L.Dropout(0.5)(x))
    v = np.ones((1, 4))
    assert np.array_equal(m.predict(v), v)                       # inference: identity
    _engine._STATE['dropout_hook'] = lambda shape, level: np.array([[1, 0, 1, 0]], dtype=bool)
    try:
        out = m._forward([v], True)[0].numpy()                   # training: kept values scaled by 1 / (1 - rate)
    finally:
        _engine._STATE['dropout_hook'] = None
    assert np.array_equal(out, [[2.0, 0.0, 2.0, 0.0]])


def test_minmaxnorm_constraint_applied_after_the_update():
    """constraints.py MinMaxNorm(0, 1) on a scalar weight: w * clip(|w|, 0, 1) / (1e-7 + |w|)"""
    c = keras.constraints.MinMaxNorm(0.0, 1.0)
    assert abs(float(c(t([1.5]))[0]) - 1.5 * 1.0 / (1e-7 + 1.5)) < 1e-15
    assert abs(float(c(t([0.4]))[0]) - 0.4 * 0.4 / (1e-7 + 0.4)) < 1e-15


def test_model_json_round_trip_with_a_lambda():
    """Model.to_json / models.model_from_json: class names, configs, connectivity; a python lambda travels as marshalled
    bytecode (utils/generic_utils.func_dump / func_load) and resolves names through custom_objects"""
    from oracle.keras_shim.keras import models
    x = keras.Input((3,), name='inp')
    d = L.Dense(2, name='d')
    y = L.Lambda(lambda z: K.exp(z) * 2.0, name='lam')(d(x))
    m = keras.Model(x, y, name='m')
    m2 = models.model_from_json(m.to_json(), {'K': K})
    m2.set_weights(m.get_weights())
    v = np.array([[0.1, -0.2, 0.3]])
    assert [l.name for l in m2.layers] == ['inp', 'd', 'lam'] and np.allclose(m2.predict(v), m.predict(v))
