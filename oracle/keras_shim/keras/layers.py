"""keras.layers (2.2.4) — the layers the LSTUR path of the reference instantiates.  TEST INFRASTRUCTURE
(oracle/keras_shim/README.md).  Each class names the Keras source file whose published behaviour it restates."""
import numpy as np
import torch

from . import activations, backend as K, constraints, initializers, regularizers
from ._engine import Input, InputLayer, Layer, Network, Sym, _has_arg, unwrap          # noqa: F401


class Dense(Layer):
    """layers/core.py: output = activation(dot(input, kernel) + bias); passes masks through"""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer='glorot_uniform', bias_initializer='zeros',
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        if 'input_shape' not in kwargs and 'input_dim' in kwargs:
            kwargs['input_shape'] = (kwargs.pop('input_dim'),)
        super().__init__(**kwargs)
        self.units, self.activation, self.use_bias = int(units), activations.get(activation), use_bias
        self.kernel_initializer, self.bias_initializer = initializers.get(kernel_initializer), initializers.get(bias_initializer)
        self.kernel_constraint, self.bias_constraint = constraints.get(kernel_constraint), constraints.get(bias_constraint)
        regularizers.get(kernel_regularizer), regularizers.get(bias_regularizer)
        self.supports_masking = True

    def build(self, input_shape):
        self.kernel = self.add_weight('kernel', (input_shape[-1], self.units), initializer=self.kernel_initializer,
                                      constraint=self.kernel_constraint)
        self.bias = self.add_weight('bias', (self.units,), initializer=self.bias_initializer,
                                    constraint=self.bias_constraint) if self.use_bias else None
        self.built = True

    def call(self, inputs):
        out = K.dot(inputs, self.kernel)
        if self.use_bias:
            out = K.bias_add(out, self.bias)
        return self.activation(out)

    def get_config(self):
        return dict(super().get_config(), units=self.units, activation=activations.serialize(self.activation), use_bias=self.use_bias)


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.supports_masking = True
        self.activation = activations.get(activation)

    def call(self, inputs):
        return self.activation(unwrap(inputs))

    def get_config(self):
        return dict(super().get_config(), activation=activations.serialize(self.activation))


class Softmax(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.supports_masking, self.axis = True, axis

    def call(self, inputs):
        return K.softmax(inputs, self.axis)


class Dropout(Layer):
    """layers/core.py: K.in_train_phase(K.dropout(inputs, rate, noise_shape), inputs); passes masks through"""

    def __init__(self, rate, noise_shape=None, seed=None, **kwargs):
        super().__init__(**kwargs)
        self.rate, self.noise_shape, self.seed = min(1., max(0., rate)), noise_shape, seed
        self.supports_masking = True

    def call(self, inputs, training=None):
        if 0. < self.rate < 1.:
            return K.in_train_phase(lambda: K.dropout(inputs, self.rate, self.noise_shape, self.seed), inputs, training=training)
        return inputs

    def get_config(self):
        return dict(super().get_config(), rate=self.rate, noise_shape=self.noise_shape)


class Masking(Layer):
    """layers/core.py: a timestep is masked when ALL its features equal mask_value; masked steps are zeroed"""

    def __init__(self, mask_value=0., **kwargs):
        super().__init__(**kwargs)
        self.supports_masking, self.mask_value = True, mask_value

    def compute_mask(self, inputs, mask=None):
        return K.any(K.not_equal(inputs, self.mask_value), axis=-1)

    def call(self, inputs):
        keep = K.any(K.not_equal(inputs, self.mask_value), axis=-1, keepdims=True)
        return inputs * K.cast(keep, K.dtype(inputs))

    def get_config(self):
        return dict(super().get_config(), mask_value=self.mask_value)


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        super().__init__(**kwargs)
        self.target_shape = tuple(target_shape)

    def call(self, inputs):
        x = unwrap(inputs)
        return x.reshape((x.shape[0],) + tuple(int(s) for s in self.target_shape))

    def get_config(self):
        return dict(super().get_config(), target_shape=self.target_shape)


class Flatten(Layer):
    def call(self, inputs):
        x = unwrap(inputs)
        return x.reshape(x.shape[0], -1)


class Lambda(Layer):
    """layers/core.py: compute_mask returns the `mask` argument (None by default): a Lambda silently drops masks"""

    def __init__(self, function, output_shape=None, mask=None, arguments=None, **kwargs):
        super().__init__(**kwargs)
        self.function, self.arguments = function, arguments or {}
        self.supports_masking = mask is not None
        self.mask = mask

    def call(self, inputs, mask=None):
        kw = dict(self.arguments)
        if _has_arg(self.function, 'mask'):
            kw['mask'] = mask
        return self.function(inputs, **kw)

    def compute_mask(self, inputs, mask=None):
        return self.mask(inputs, mask) if callable(self.mask) else self.mask

    def get_config(self):
        # layers/core.py: a python lambda travels as its marshalled bytecode (utils/generic_utils.func_dump), a named function
        # by name; only readable by the same python version — true of Keras too
        if getattr(self.function, '__name__', '') == '<lambda>':
            function, function_type = func_dump(self.function), 'lambda'
        else:
            function, function_type = self.function.__name__, 'function'
        return dict(super().get_config(), function=function, function_type=function_type, output_shape=None,
                    output_shape_type='raw', arguments=self.arguments)

    @classmethod
    def from_config(cls, config, custom_objects=None):
        config = dict(config)
        globs = dict(globals())
        globs.update(custom_objects or {})
        ftype = config.pop('function_type')
        if ftype == 'lambda':
            function = func_load(config['function'], globs=globs)
        else:
            function = globs[config['function']]
        return cls(function, arguments=config.get('arguments') or {}, name=config.get('name'))


def func_dump(func):
    """utils/generic_utils.func_dump: (base64 of the marshalled code object, defaults, closure cell contents)"""
    import codecs
    import marshal
    code = codecs.encode(marshal.dumps(func.__code__), 'base64').decode('ascii')
    closure = tuple(c.cell_contents for c in func.__closure__) if func.__closure__ else None
    return code, func.__defaults__, closure


def func_load(code, defaults=None, closure=None, globs=None):
    """utils/generic_utils.func_load"""
    import codecs
    import marshal
    import types
    if isinstance(code, (tuple, list)):
        code, defaults, closure = code
        if isinstance(defaults, list):
            defaults = tuple(defaults)

    def ensure_value_to_cell(value):
        def dummy_fn():
            value          # noqa: B018  (makes `value` a closure cell)
        cell = dummy_fn.__closure__[0]
        return value if isinstance(value, type(cell)) else cell

    if closure is not None:
        closure = tuple(ensure_value_to_cell(v) for v in closure)
    raw = codecs.decode(code.encode('ascii'), 'base64')
    return types.FunctionType(marshal.loads(raw), globs if globs is not None else globals(), name='<lambda>',
                              argdefs=defaults, closure=closure)


class Embedding(Layer):
    """layers/embeddings.py: K.gather(embeddings, int32(inputs)); mask = (inputs != 0) only with mask_zero"""

    def __init__(self, input_dim, output_dim, embeddings_initializer='uniform', embeddings_regularizer=None,
                 activity_regularizer=None, embeddings_constraint=None, mask_zero=False, input_length=None, **kwargs):
        if 'input_shape' not in kwargs:
            kwargs['input_shape'] = (input_length,) if input_length else (None,)
        super().__init__(**kwargs)
        self.input_dim, self.output_dim, self.mask_zero, self.input_length = int(input_dim), int(output_dim), mask_zero, input_length
        self.embeddings_initializer = initializers.get(embeddings_initializer)
        self.embeddings_constraint = constraints.get(embeddings_constraint)
        self.supports_masking = mask_zero

    def build(self, input_shape):
        self.embeddings = self.add_weight('embeddings', (self.input_dim, self.output_dim), initializer=self.embeddings_initializer,
                                          constraint=self.embeddings_constraint)
        self.built = True

    def compute_mask(self, inputs, mask=None):
        return K.not_equal(inputs, 0) if self.mask_zero else None

    def call(self, inputs):
        return K.gather(self.embeddings, K.cast(inputs, 'int32'))

    def get_config(self):
        return dict(super().get_config(), input_dim=self.input_dim, output_dim=self.output_dim, mask_zero=self.mask_zero,
                    input_length=self.input_length)


class Conv1D(Layer):
    """layers/convolutional.py: kernel (k, in, out); 'same' pads (k-1)//2 left, the rest right (TensorFlow SAME, stride 1)"""

    def __init__(self, filters, kernel_size, strides=1, padding='valid', data_format='channels_last', dilation_rate=1,
                 activation=None, use_bias=True, kernel_initializer='glorot_uniform', bias_initializer='zeros', **kwargs):
        for k in ('kernel_regularizer', 'bias_regularizer', 'activity_regularizer', 'kernel_constraint', 'bias_constraint'):
            kwargs.pop(k, None)
        super().__init__(**kwargs)
        self.filters = int(filters)
        self.kernel_size = int(kernel_size[0] if isinstance(kernel_size, (list, tuple)) else kernel_size)
        self.strides = int(strides[0] if isinstance(strides, (list, tuple)) else strides)
        self.padding, self.activation, self.use_bias = padding, activations.get(activation), use_bias
        self.kernel_initializer, self.bias_initializer = initializers.get(kernel_initializer), initializers.get(bias_initializer)
        if self.strides != 1 or dilation_rate not in (1, (1,), [1]) or padding not in ('same', 'valid'):
            raise NotImplementedError('Conv1D: only stride 1, dilation 1, same / valid')

    def build(self, input_shape):
        self.kernel = self.add_weight('kernel', (self.kernel_size, input_shape[-1], self.filters), initializer=self.kernel_initializer)
        self.bias = self.add_weight('bias', (self.filters,), initializer=self.bias_initializer) if self.use_bias else None
        self.built = True

    def call(self, inputs):
        x = unwrap(inputs).transpose(1, 2)                   # (B, in, T)
        if self.padding == 'same':
            left = (self.kernel_size - 1) // 2
            x = torch.nn.functional.pad(x, (left, self.kernel_size - 1 - left))
        out = torch.nn.functional.conv1d(x, self.kernel.permute(2, 1, 0)).transpose(1, 2)
        if self.use_bias:
            out = out + self.bias
        return self.activation(out)

    def get_config(self):
        return dict(super().get_config(), filters=self.filters, kernel_size=(self.kernel_size,), padding=self.padding,
                    activation=activations.serialize(self.activation), use_bias=self.use_bias)


Convolution1D = Conv1D


# ------------------------------------------------------------------------------------------------ merge layers
class _Merge(Layer):
    """layers/merge.py: supports_masking = True"""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.supports_masking = True

    def compute_mask(self, inputs, mask=None):
        if mask is None or all(m is None for m in mask):
            return None
        masks = [K.expand_dims(m, 0) for m in mask if m is not None]
        return K.all(K.concatenate(masks, axis=0), axis=0, keepdims=False)


class Add(_Merge):
    def call(self, inputs):
        out = unwrap(inputs[0])
        for x in inputs[1:]:
            out = out + unwrap(x)
        return out


class Multiply(_Merge):
    def call(self, inputs):
        out = unwrap(inputs[0])
        for x in inputs[1:]:
            out = out * unwrap(x)
        return out


class Concatenate(_Merge):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, inputs):
        return K.concatenate(list(inputs), axis=self.axis)

    def compute_mask(self, inputs, mask=None):
        if mask is None or all(m is None for m in mask):
            return None
        masks = []
        for x, m in zip(inputs, mask):
            x = unwrap(x)
            if m is None:
                masks.append(torch.ones_like(x, dtype=torch.bool))
            elif m.dim() < x.dim():
                masks.append(K.expand_dims(m))
            else:
                masks.append(m)
        return K.all(K.concatenate(masks, axis=self.axis), axis=-1, keepdims=False)

    def get_config(self):
        return dict(super().get_config(), axis=self.axis)


class Dot(_Merge):
    def __init__(self, axes, normalize=False, **kwargs):
        super().__init__(**kwargs)
        self.axes, self.normalize = axes, normalize

    def build(self, input_shape):
        # layers/merge.py Dot.build: the contracted dimensions must agree
        s1, s2 = input_shape[0], input_shape[1]
        if isinstance(self.axes, int):
            axes = [self.axes % len(s1), self.axes % len(s2)] if self.axes < 0 else [self.axes] * 2
        else:
            axes = list(self.axes)
        if s1[axes[0]] != s2[axes[1]]:
            raise ValueError('Dimension incompatibility %s != %s. Layer shapes: %s, %s' % (s1[axes[0]], s2[axes[1]], s1, s2))
        self.built = True

    def call(self, inputs):
        x1, x2 = unwrap(inputs[0]), unwrap(inputs[1])
        if isinstance(self.axes, int):
            axes = [self.axes % x1.dim(), self.axes % x2.dim()] if self.axes < 0 else [self.axes] * 2
        else:
            axes = [self.axes[i] % (x1, x2)[i].dim() if self.axes[i] < 0 else self.axes[i] for i in range(2)]
        if self.normalize:
            x1 = x1 / torch.sqrt(torch.clamp((x1 ** 2).sum(axes[0], keepdim=True), min=K.epsilon()))
            x2 = x2 / torch.sqrt(torch.clamp((x2 ** 2).sum(axes[1], keepdim=True), min=K.epsilon()))
        return K.batch_dot(x1, x2, axes)

    def compute_mask(self, inputs, mask=None):
        return None

    def get_config(self):
        return dict(super().get_config(), axes=self.axes, normalize=self.normalize)


def add(inputs, **kwargs):
    return Add(**kwargs)(inputs)


def multiply(inputs, **kwargs):
    return Multiply(**kwargs)(inputs)


def concatenate(inputs, axis=-1, **kwargs):
    return Concatenate(axis=axis, **kwargs)(inputs)


def dot(inputs, axes, normalize=False, **kwargs):
    return Dot(axes=axes, normalize=normalize, **kwargs)(inputs)


# ------------------------------------------------------------------------------------------------ wrappers
class Wrapper(Layer):
    def __init__(self, layer, **kwargs):
        self.layer = layer
        super().__init__(**kwargs)

    @property
    def trainable_weights(self):
        return self.layer.trainable_weights if self.trainable else []

    @property
    def non_trainable_weights(self):
        return self.layer.non_trainable_weights if self.trainable else self.layer.weights

    @property
    def weights(self):
        return self.trainable_weights + self.non_trainable_weights

    def get_config(self):
        return dict(super().get_config(), layer={'class_name': self.layer.__class__.__name__, 'config': self.layer.get_config()})


class TimeDistributed(Wrapper):
    """layers/wrappers.py (batch size unknown): fold time into the batch, apply the layer, unfold"""

    def __init__(self, layer, **kwargs):
        super().__init__(layer, **kwargs)
        self.supports_masking = True

    def build(self, input_shape):
        if not self.layer.built:
            self.layer.build((input_shape[0],) + tuple(input_shape[2:]))
            self.layer.built = True
        self.built = True

    def call(self, inputs, training=None, mask=None):
        x = unwrap(inputs)
        b, t = x.shape[0], x.shape[1]
        flat = x.reshape((b * t,) + tuple(x.shape[2:]))
        kw = {'training': training} if _has_arg(self.layer.call, 'training') else {}
        y = self.layer.call(flat, **kw)
        return y.reshape((b, t) + tuple(y.shape[1:]))

    def compute_mask(self, inputs, mask=None):
        return mask          # 2.2.4: derived from the input mask (None here: the history input carries none)


# ------------------------------------------------------------------------------------------------ recurrent
class RNN(Layer):
    """layers/recurrent.py.  K.rnn with a mask: a masked step keeps the previous state AND re-emits the previous output
    (states[0]); the initial previous output is zeros.  `initial_state` tensors join the inputs of the node."""
    n_states = 1

    def __init__(self, units, return_sequences=False, return_state=False, go_backwards=False, stateful=False, unroll=False, **kwargs):
        super().__init__(**kwargs)
        self.units, self.return_sequences, self.return_state, self.go_backwards = int(units), return_sequences, return_state, go_backwards
        self.supports_masking = True

    def __call__(self, inputs, initial_state=None, **kwargs):
        if initial_state is not None:
            if isinstance(inputs, (list, tuple)):
                raise ValueError('initial_state given twice')
            st = list(initial_state) if isinstance(initial_state, (list, tuple)) else [initial_state]
            return super().__call__([inputs] + st, **kwargs)
        return super().__call__(inputs, **kwargs)

    def _maybe_build(self, xs, list_in):
        return super()._maybe_build(xs[:1], False)

    def compute_mask(self, inputs, mask=None):
        if isinstance(mask, (list, tuple)):
            mask = mask[0]
        out = mask if self.return_sequences else None
        return [out] + [None] * self.n_states if self.return_state else out

    def call(self, inputs, mask=None, training=None, initial_state=None):
        if isinstance(inputs, (list, tuple)):
            initial_state, inputs = list(inputs[1:]), inputs[0]
        if isinstance(mask, (list, tuple)):
            mask = mask[0]
        x = unwrap(inputs)
        b, t = x.shape[0], x.shape[1]
        if not initial_state:
            states = [torch.zeros(b, self.units, dtype=x.dtype) for _ in range(self.n_states)]
        else:
            states = [unwrap(s) for s in initial_state]
            if len(states) != self.n_states:
                raise ValueError('Layer has %d states but was passed %d initial states' % (self.n_states, len(states)))
        steps = range(t - 1, -1, -1) if self.go_backwards else range(t)
        outputs = []
        for i in steps:
            out, new_states = self.step(x[:, i], states)
            if mask is not None:
                m = mask[:, i].bool().unsqueeze(-1)
                out = torch.where(m, out, states[0])         # tf.where(tiled_mask_t, output, states[0])
                new_states = [torch.where(m, n, s) for n, s in zip(new_states, states)]
            outputs.append(out)
            states = new_states
        y = torch.stack(outputs, 1) if self.return_sequences else outputs[-1]
        return [y] + list(states) if self.return_state else y


class GRU(RNN):
    """GRUCell, implementation 1, reset_after=False: z, r, h gate order; h_t = z * h + (1 - z) * act(x_h + (r * h) U_h)"""

    def __init__(self, units, activation='tanh', recurrent_activation='hard_sigmoid', use_bias=True,
                 kernel_initializer='glorot_uniform', recurrent_initializer='orthogonal', bias_initializer='zeros',
                 dropout=0., recurrent_dropout=0., implementation=1, reset_after=False, **kwargs):
        rnn_kw = {k: kwargs.pop(k) for k in ('return_sequences', 'return_state', 'go_backwards', 'stateful', 'unroll') if k in kwargs}
        super().__init__(units, **rnn_kw, **kwargs)
        if dropout or recurrent_dropout or reset_after:
            raise NotImplementedError('GRU: dropout / reset_after are not used by the LSTUR path')
        self.activation, self.recurrent_activation = activations.get(activation), activations.get(recurrent_activation)
        self.use_bias = use_bias
        self.kernel_initializer, self.recurrent_initializer = initializers.get(kernel_initializer), initializers.get(recurrent_initializer)
        self.bias_initializer = initializers.get(bias_initializer)

    def build(self, input_shape):
        u = self.units
        self.kernel = self.add_weight('kernel', (input_shape[-1], 3 * u), initializer=self.kernel_initializer)
        self.recurrent_kernel = self.add_weight('recurrent_kernel', (u, 3 * u), initializer=self.recurrent_initializer)
        self.bias = self.add_weight('bias', (3 * u,), initializer=self.bias_initializer) if self.use_bias else None
        self.built = True

    def step(self, x, states):
        h, u = states[0], self.units
        xw = x @ self.kernel
        if self.use_bias:
            xw = xw + self.bias
        rk = self.recurrent_kernel
        z = self.recurrent_activation(xw[:, :u] + h @ rk[:, :u])
        r = self.recurrent_activation(xw[:, u:2 * u] + h @ rk[:, u:2 * u])
        hh = self.activation(xw[:, 2 * u:] + (r * h) @ rk[:, 2 * u:])
        h = z * h + (1 - z) * hh
        return h, [h]

    def get_config(self):
        return dict(super().get_config(), units=self.units, return_sequences=self.return_sequences,
                    activation=activations.serialize(self.activation),
                    recurrent_activation=activations.serialize(self.recurrent_activation))


class LSTM(RNN):
    """LSTMCell, implementation 1: i, f, c, o gate order, unit_forget_bias"""
    n_states = 2

    def __init__(self, units, activation='tanh', recurrent_activation='hard_sigmoid', use_bias=True,
                 kernel_initializer='glorot_uniform', recurrent_initializer='orthogonal', bias_initializer='zeros',
                 unit_forget_bias=True, dropout=0., recurrent_dropout=0., implementation=1, **kwargs):
        rnn_kw = {k: kwargs.pop(k) for k in ('return_sequences', 'return_state', 'go_backwards', 'stateful', 'unroll') if k in kwargs}
        super().__init__(units, **rnn_kw, **kwargs)
        if dropout or recurrent_dropout:
            raise NotImplementedError('LSTM: dropout is not used by the LSTUR path')
        self.activation, self.recurrent_activation = activations.get(activation), activations.get(recurrent_activation)
        self.use_bias, self.unit_forget_bias = use_bias, unit_forget_bias
        self.kernel_initializer, self.recurrent_initializer = initializers.get(kernel_initializer), initializers.get(recurrent_initializer)
        self.bias_initializer = initializers.get(bias_initializer)

    def build(self, input_shape):
        u = self.units
        self.kernel = self.add_weight('kernel', (input_shape[-1], 4 * u), initializer=self.kernel_initializer)
        self.recurrent_kernel = self.add_weight('recurrent_kernel', (u, 4 * u), initializer=self.recurrent_initializer)
        if self.use_bias:
            if self.unit_forget_bias:
                init = lambda shape, dtype=None: np.concatenate([self.bias_initializer((u,)), np.ones(u), self.bias_initializer((2 * u,))])
            else:
                init = self.bias_initializer
            self.bias = self.add_weight('bias', (4 * u,), initializer=init)
        else:
            self.bias = None
        self.built = True

    def step(self, x, states):
        h, c, u = states[0], states[1], self.units
        xw = x @ self.kernel
        if self.use_bias:
            xw = xw + self.bias
        hw = h @ self.recurrent_kernel
        i = self.recurrent_activation(xw[:, :u] + hw[:, :u])
        f = self.recurrent_activation(xw[:, u:2 * u] + hw[:, u:2 * u])
        c = f * c + i * self.activation(xw[:, 2 * u:3 * u] + hw[:, 2 * u:3 * u])
        o = self.recurrent_activation(xw[:, 3 * u:] + hw[:, 3 * u:])
        h = o * self.activation(c)
        return h, [h, c]


# ------------------------------------------------------------------------------------------------ named but unused on this path
def _unused(name):
    class _Unused(Layer):
        def __init__(self, *a, **kw):
            kw = {k: v for k, v in kw.items() if k in ('name', 'trainable')}
            super().__init__(**kw)

        def call(self, inputs, **kwargs):
            raise NotImplementedError('keras.layers.%s is not on the LSTUR path (oracle/keras_shim/README.md)' % name)
    _Unused.__name__ = name
    return _Unused


for _n in ('BatchNormalization', 'Bidirectional', 'GlobalMaxPooling1D', 'GlobalAveragePooling1D', 'MaxPooling1D', 'Conv2D',
           'SimpleRNN', 'RepeatVector', 'Permute', 'Average', 'Maximum', 'Subtract', 'LeakyReLU', 'AveragePooling1D',
           'ZeroPadding1D', 'SpatialDropout1D', 'GaussianNoise', 'CuDNNGRU', 'CuDNNLSTM', 'Cropping1D'):
    globals()[_n] = _unused(_n)
