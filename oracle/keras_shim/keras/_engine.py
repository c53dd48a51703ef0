"""Core of the Keras-2.2 restatement (see ../README.md): symbolic tensors, layers, functional Model, training loop.

TEST INFRASTRUCTURE ONLY.  Semantics follow stand-alone Keras 2.2.4 with the TensorFlow-1.x backend, file by file:
  engine/base_layer.py   Layer.__call__ (build on first call, mask collection, compute_mask contract)
  engine/network.py      Network._init_graph_network (layer order by depth), run_internal_graph
  engine/training.py     compile / train_on_batch / predict / fit_generator (loss = mean over the batch)
  layers/*.py            the layers listed in ../README.md
  optimizers.py          Adam, SGD
Graphs are evaluated eagerly with torch; a symbolic tensor carries a small EXAMPLE value (batch of 2) that stands in for
shape inference, so Lambda layers and custom `call`s need no `compute_output_shape`.
"""
import inspect
import re

import numpy as np
import torch

# ------------------------------------------------------------------------------------------------ backend state
_STATE = {'floatx': 'float32', 'epsilon': 1e-7, 'phase': 0, 'ctx': None, 'dropout_hook': None, 'uids': {}}
EXAMPLE_BATCH = 2


def floatx():
    return _STATE['floatx']


def tdtype(name=None):
    name = name or floatx()
    if isinstance(name, torch.dtype):
        return name
    return {'float32': torch.float32, 'float64': torch.float64, 'float16': torch.float16, 'int32': torch.int32,
            'int64': torch.int64, 'bool': torch.bool, 'float': torch.float32}[str(name)]


def get_uid(prefix):
    _STATE['uids'][prefix] = _STATE['uids'].get(prefix, 0) + 1
    return _STATE['uids'][prefix]


def reset_uids():
    _STATE['uids'].clear()


def to_snake_case(name):
    s = re.sub('(.)([A-Z][a-z0-9]+)', r'\1_\2', name)
    s = re.sub('([a-z])([A-Z])', r'\1_\2', s).lower()
    return 'private' + s if s[0] == '_' else s


# ------------------------------------------------------------------------------------------------ symbolic tensors
class Dimension:
    def __init__(self, value):
        self.value = value

    def __int__(self):
        return int(self.value)

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return 'Dimension(%r)' % (self.value,)


class TensorShape:
    def __init__(self, dims):
        self.dims = [d if isinstance(d, Dimension) else Dimension(d) for d in dims]

    def __getitem__(self, i):
        return TensorShape(self.dims[i]) if isinstance(i, slice) else self.dims[i]

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(self.dims)

    def as_list(self):
        return [d.value for d in self.dims]

    @property
    def ndims(self):
        return len(self.dims)


class Ctx:
    """values of the symbolic tensors of ONE evaluation of a graph; nested models chain to the caller's context"""

    def __init__(self, parent=None, training=False):
        self.memo, self.parent, self.training = {}, parent, training

    def find(self, sym):
        c = self
        while c is not None:
            if id(sym) in c.memo:
                return c.memo[id(sym)]
            c = c.parent
        return None


class Sym:
    """symbolic tensor: output `idx` of `node`; `example` = the value it takes for an EXAMPLE_BATCH example input"""

    def __init__(self, node, idx, example, mask_example=None, name=None):
        self.node, self.idx, self.example, self.mask_example, self.name = node, idx, example, mask_example, name
        self._keras_shape = (None,) + tuple(example.shape[1:])
        self._uses_learning_phase = False

    @property
    def shape(self):
        return TensorShape(self._keras_shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return str(self.example.dtype).replace('torch.', '')

    @property
    def _keras_history(self):
        return (self.node.layer, self.node.layer._inbound_nodes.index(self.node), self.idx)

    def value(self):
        return eval_sym(self)[0]

    # a symbolic tensor captured by a custom layer (models.GlobalAveragePoolingMasked(mask)) and used inside its call()
    def __mul__(self, o):
        return self.value() * unwrap(o)

    __rmul__ = __mul__

    def __add__(self, o):
        return self.value() + unwrap(o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.value() - unwrap(o)

    def __rsub__(self, o):
        return unwrap(o) - self.value()

    def __truediv__(self, o):
        return self.value() / unwrap(o)

    def __rtruediv__(self, o):
        return unwrap(o) / self.value()

    def __getitem__(self, k):
        return self.value()[k]

    def __repr__(self):
        return '<Sym %s %s of %s>' % (self.name or '', self._keras_shape, self.node.layer.name)


def unwrap(x):
    """Sym -> its value in the current evaluation; python / numpy scalars and arrays -> tensors"""
    if isinstance(x, Sym):
        return x.value()
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)):
        return type(x)(unwrap(v) for v in x)
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x)
    return x


def eval_sym(sym):
    """(value, mask) of a symbolic tensor in the current context, evaluating its producers on demand"""
    ctx = _STATE['ctx']
    if ctx is None:             # graph-construction time: the example value
        return sym.example, sym.mask_example
    hit = ctx.find(sym)
    if hit is not None:
        return hit
    node = sym.node
    if isinstance(node.layer, InputLayer):
        if _STATE.get('building', 0):       # example evaluation while the graph is being assembled
            return sym.example, sym.mask_example
        raise RuntimeError('input %r is not fed' % (sym,))
    ins = [eval_sym(s) for s in node.inputs]
    outs, masks = node.layer._run([v for v, _ in ins], [m for _, m in ins], node.list_in, node.kwargs, ctx.training)
    for s, v, m in zip(node.outputs, outs, masks):
        ctx.memo[id(s)] = (v, m)
    return ctx.memo[id(sym)]


class Node:
    def __init__(self, layer, inputs, list_in, kwargs):
        self.layer, self.inputs, self.list_in, self.kwargs, self.outputs = layer, inputs, list_in, kwargs, []
        self.id = get_uid('__node__')

    @property
    def inbound_layers(self):
        return [s.node.layer for s in self.inputs]

    @property
    def inbound_nodes(self):
        return [s.node for s in self.inputs]


# ------------------------------------------------------------------------------------------------ variables
class Variable(torch.nn.Parameter):
    """a weight: torch Parameter + Keras' name / constraint"""

    def __new__(cls, data, name=None, constraint=None, trainable=True):
        v = torch.nn.Parameter.__new__(cls, data, requires_grad=True)
        v.vname, v.constraint, v.vtrainable = name, constraint, trainable
        return v

    def __deepcopy__(self, memo):
        return Variable(self.data.clone(), self.vname, self.constraint, self.vtrainable)


def variable(value, dtype=None, name=None, constraint=None):
    return Variable(torch.as_tensor(np.asarray(value), dtype=tdtype(dtype)).clone(), name, constraint)


# ------------------------------------------------------------------------------------------------ Layer
def _has_arg(fn, name):
    try:
        return name in inspect.signature(fn).parameters
    except (TypeError, ValueError):
        return False


def _is_all_none(m):
    return m is None or (isinstance(m, (list, tuple)) and all(x is None for x in m))


class Layer:
    def __init__(self, **kwargs):
        allowed = {'input_shape', 'batch_input_shape', 'batch_size', 'dtype', 'name', 'trainable', 'weights', 'input_dtype',
                   'input_dim', 'input_length'}
        for k in kwargs:
            if k not in allowed:
                raise TypeError('Keyword argument not understood: ' + k)
        name = kwargs.get('name')
        if not name:
            prefix = to_snake_case(self.__class__.__name__)
            name = prefix + '_' + str(get_uid(prefix))
        self.name = name
        self.trainable = kwargs.get('trainable', True)
        self._initial_weights = kwargs.get('weights')
        self.dtype = kwargs.get('dtype') or floatx()
        if not hasattr(self, 'supports_masking'):
            self.supports_masking = False
        self.built = False
        self._trainable_weights, self._non_trainable_weights = [], []
        self._inbound_nodes = []
        self.stateful = False
        self.input_spec = None

    # -- weights
    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True, constraint=None):
        from . import initializers
        init = initializers.get(initializer if initializer is not None else 'glorot_uniform')
        v = Variable(torch.as_tensor(np.asarray(init(tuple(int(s) for s in shape))), dtype=tdtype(dtype or self.dtype)).clone(),
                     name=self.name + '/' + str(name), constraint=constraint, trainable=trainable)
        (self._trainable_weights if trainable else self._non_trainable_weights).append(v)
        return v

    @property
    def trainable_weights(self):
        return list(self._trainable_weights) if self.trainable else []

    @property
    def non_trainable_weights(self):
        return list(self._non_trainable_weights) if self.trainable else list(self._trainable_weights) + list(self._non_trainable_weights)

    @property
    def weights(self):
        return self.trainable_weights + self.non_trainable_weights

    def get_weights(self):
        return [w.detach().cpu().numpy().copy() for w in self.weights]

    def set_weights(self, weights):
        ws = self.weights
        if len(ws) != len(weights):
            raise ValueError('layer %s: expected %d weight arrays, got %d' % (self.name, len(ws), len(weights)))
        for w, a in zip(ws, weights):
            a = np.asarray(a)
            if tuple(w.shape) != tuple(a.shape):
                raise ValueError('layer %s: weight shape %s vs %s' % (self.name, tuple(w.shape), a.shape))
            w.data.copy_(torch.as_tensor(a, dtype=w.dtype))

    # -- the contract of engine/base_layer.py
    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        return inputs

    def compute_mask(self, inputs, mask=None):
        if not self.supports_masking:
            if mask is not None:
                if isinstance(mask, (list, tuple)):
                    if any(m is not None for m in mask):
                        raise TypeError('Layer ' + self.name + ' does not support masking, but was passed an input_mask')
                else:
                    raise TypeError('Layer ' + self.name + ' does not support masking, but was passed an input_mask')
            return None
        return mask

    def compute_output_shape(self, input_shape):
        return input_shape

    def get_config(self):
        return {'name': self.name, 'trainable': self.trainable}

    def _shape_of(self, x):
        return x._keras_shape if isinstance(x, Sym) else (None,) + tuple(x.shape[1:])

    def _maybe_build(self, xs, list_in):
        if not self.built:
            shapes = [self._shape_of(x) for x in xs]
            self.build(shapes if list_in else shapes[0])
            self.built = True
            if self._initial_weights is not None:
                self.set_weights(self._initial_weights)
                self._initial_weights = None

    def _run(self, values, masks, list_in, kwargs, training):
        """one evaluation: call() + compute_mask() as Layer.__call__ wires them"""
        x = list(values) if list_in else values[0]
        m = list(masks) if list_in else masks[0]
        kw = dict(kwargs)
        if not _is_all_none(m) and _has_arg(self.call, 'mask') and 'mask' not in kw:
            kw['mask'] = m
        if _has_arg(self.call, 'training') and 'training' not in kw:
            kw['training'] = training
        prev_phase = _STATE['phase']
        _STATE['phase'] = 1 if training else 0
        try:
            out = self.call(x, **kw)
        finally:
            _STATE['phase'] = prev_phase
        out_mask = self.compute_mask(x, m)
        outs = list(out) if isinstance(out, (list, tuple)) else [out]
        if isinstance(out_mask, (list, tuple)):
            out_masks = list(out_mask)
        else:
            out_masks = [out_mask] * len(outs)
        return outs, out_masks

    def __call__(self, inputs, **kwargs):
        list_in = isinstance(inputs, (list, tuple))
        xs = list(inputs) if list_in else [inputs]
        symbolic = any(isinstance(x, Sym) for x in xs)
        self._maybe_build(xs, list_in)
        if not symbolic:        # a layer applied to VALUES inside another layer's call(): plain evaluation
            outs, _ = self._run([unwrap(x) for x in xs], [None] * len(xs), list_in, kwargs, bool(_STATE['phase']))
            return outs if len(outs) > 1 else outs[0]
        if not all(isinstance(x, Sym) for x in xs):
            raise ValueError('layer %s called with a mix of symbolic and concrete inputs' % self.name)
        node = Node(self, xs, list_in, kwargs)
        saved = _STATE['ctx']
        _STATE['ctx'] = None
        _STATE['building'] = _STATE.get('building', 0) + 1
        try:
            with torch.no_grad():
                outs, masks = self._run([x.example for x in xs], [x.mask_example for x in xs], list_in, kwargs, False)
        finally:
            _STATE['ctx'] = saved
            _STATE['building'] -= 1
        node.outputs = [Sym(node, i, o, m) for i, (o, m) in enumerate(zip(outs, masks))]
        self._inbound_nodes.append(node)
        return node.outputs if len(node.outputs) > 1 else node.outputs[0]

    # Keras' accessors used by the reference (task/test_pipeline.py: get_layer(...).input / .output)
    @property
    def input(self):
        n = self._inbound_nodes[0]
        return n.inputs if n.list_in else n.inputs[0]

    @property
    def output(self):
        n = self._inbound_nodes[0]
        return n.outputs if len(n.outputs) > 1 else n.outputs[0]

    def get_input_at(self, i):
        n = self._inbound_nodes[i]
        return n.inputs if n.list_in else n.inputs[0]

    def get_output_at(self, i):
        n = self._inbound_nodes[i]
        return n.outputs if len(n.outputs) > 1 else n.outputs[0]

    @property
    def input_shape(self):
        x = self.input
        return [s._keras_shape for s in x] if isinstance(x, list) else x._keras_shape

    @property
    def output_shape(self):
        x = self.output
        return [s._keras_shape for s in x] if isinstance(x, list) else x._keras_shape


class InputLayer(Layer):
    def __init__(self, shape, dtype=None, name=None):
        if not name:
            name = 'input_' + str(get_uid('input'))
        super().__init__(name=name, dtype=dtype or floatx())
        self.built = True
        self.batch_input_shape = (None,) + tuple(shape)
        node = Node(self, [], False, {})
        ex = torch.zeros((EXAMPLE_BATCH,) + tuple(int(s) for s in shape), dtype=tdtype(self.dtype))
        node.outputs = [Sym(node, 0, ex, None, name=self.name)]
        self._inbound_nodes.append(node)

    @property
    def input_shape(self):
        return self.batch_input_shape

    @property
    def output_shape(self):
        return self.batch_input_shape

    def get_config(self):
        return {'batch_input_shape': list(self.batch_input_shape), 'dtype': self.dtype, 'sparse': False, 'name': self.name}

    @property
    def input(self):
        return self._inbound_nodes[0].outputs[0]


def Input(shape=None, batch_shape=None, name=None, dtype=None, sparse=False, tensor=None):
    if shape is None:
        shape = tuple(batch_shape[1:])
    return InputLayer(tuple(shape), dtype=dtype, name=name)._inbound_nodes[0].outputs[0]


# ------------------------------------------------------------------------------------------------ Network / Model
class Network(Layer):
    """engine/network.py: a graph of layers that is itself a layer"""

    def __init__(self, inputs, outputs, name=None):
        if not name:
            prefix = self.__class__.__name__.lower()
            name = prefix + '_' + str(get_uid(prefix))
        Layer.__init__(self, name=name)
        self.supports_masking = False
        self._list_in = isinstance(inputs, (list, tuple))
        self._list_out = isinstance(outputs, (list, tuple))
        self.inputs = list(inputs) if self._list_in else [inputs]
        self.outputs = list(outputs) if self._list_out else [outputs]
        self.built = True
        self._init_graph()

    def _init_graph(self):
        # Network._init_graph_network: depth-first walk from the outputs; layers ordered by decreasing depth, ties by
        # first visit.  This order is what model.layers / get_weights() / the pkl files of utils.save_model follow.
        nodes_in_decreasing_depth, finished, in_progress, layer_indices = [], set(), set(), {}

        def build_map(sym):
            node, layer = sym.node, sym.node.layer
            if id(node) in in_progress:
                raise ValueError('The tensor %r at layer "%s" is part of a cycle.' % (sym, layer.name))
            if id(node) in finished:
                return
            if id(layer) not in layer_indices:
                layer_indices[id(layer)] = len(layer_indices)
            in_progress.add(id(node))
            for x in node.inputs:
                build_map(x)
            finished.add(id(node))
            in_progress.discard(id(node))
            nodes_in_decreasing_depth.append(node)

        for x in self.outputs:
            build_map(x)
        nodes_depths, layers_depths, layers = {}, {}, {}
        for node in reversed(nodes_in_decreasing_depth):
            depth = nodes_depths.setdefault(id(node), 0)
            depth = max(depth, layers_depths.get(id(node.layer), 0))
            layers_depths[id(node.layer)] = depth
            layers[id(node.layer)] = node.layer
            nodes_depths[id(node)] = depth
            for s in node.inputs:
                nodes_depths[id(s.node)] = max(depth + 1, nodes_depths.get(id(s.node), 0))
        by_depth = {}
        for lid, depth in layers_depths.items():
            by_depth.setdefault(depth, []).append(layers[lid])
        self.layers = []
        for depth in sorted(by_depth, reverse=True):
            self.layers.extend(sorted(by_depth[depth], key=lambda l: layer_indices[id(l)]))
        self._nodes = nodes_in_decreasing_depth
        fed = {id(s.node) for s in self.inputs}
        for node in nodes_in_decreasing_depth:
            if isinstance(node.layer, InputLayer) and id(node) not in fed:
                raise ValueError('Graph disconnected: cannot obtain value for tensor %r at layer "%s".' % (node.outputs[0], node.layer.name))
        self.input_names = [s.node.layer.name for s in self.inputs]
        self.output_names = [s.node.layer.name for s in self.outputs]

    def summary(self, *a, **kw):
        pass

    def count_params(self):
        seen, n = set(), 0
        for w in self.weights:
            if id(w) not in seen:
                seen.add(id(w))
                n += int(w.numel())
        return n

    def get_layer(self, name=None, index=None):
        if index is not None:
            return self.layers[index]
        for layer in self.layers:
            if layer.name == name:
                return layer
        raise ValueError('No such layer: ' + str(name))

    @property
    def trainable_weights(self):
        if not self.trainable:
            return []
        out = []
        for layer in self.layers:
            out += layer.trainable_weights
        return out

    @property
    def non_trainable_weights(self):
        out = []
        for layer in self.layers:
            out += layer.non_trainable_weights
        if not self.trainable:
            tw = []
            for layer in self.layers:
                tw += layer.trainable_weights
            return tw + out
        return out

    def get_weights(self):      # Network.get_weights(): every layer's weights in layer order (shared layers repeat)
        out = []
        for layer in self.layers:
            out += layer.weights
        return [w.detach().cpu().numpy().copy() for w in out]

    def set_weights(self, weights):
        weights = list(weights)
        for layer in self.layers:
            n = len(layer.weights)
            Layer.set_weights(layer, weights[:n]) if not isinstance(layer, Network) else layer._set_flat(weights[:n])
            weights = weights[n:]

    def _set_flat(self, weights):       # a nested model's `weights` = trainable + non-trainable (Layer.weights)
        for w, a in zip(self.weights, weights):
            w.data.copy_(torch.as_tensor(np.asarray(a), dtype=w.dtype))

    def _evaluate(self, values, masks, training):
        ctx = Ctx(parent=_STATE['ctx'], training=training)
        if len(values) != len(self.inputs):
            raise ValueError('model %s expects %d inputs, got %d' % (self.name, len(self.inputs), len(values)))
        for s, v, m in zip(self.inputs, values, masks):
            ctx.memo[id(s)] = (v, m)
        saved = _STATE['ctx']
        _STATE['ctx'] = ctx
        try:
            res = [eval_sym(s) for s in self.outputs]
        finally:
            _STATE['ctx'] = saved
        return [v for v, _ in res], [m for _, m in res]

    # as a layer inside another graph
    def _run(self, values, masks, list_in, kwargs, training):
        return self._evaluate(values, masks, training)

    def call(self, inputs, mask=None, training=None):
        xs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        ms = list(mask) if isinstance(mask, (list, tuple)) else [mask] * len(xs)
        training = bool(_STATE['phase']) if training is None else training
        outs, _ = self._evaluate([unwrap(x) for x in xs], ms, training)
        return outs if self._list_out else outs[0]

    @property
    def input(self):
        return self.inputs if self._list_in else self.inputs[0]

    @property
    def output(self):
        return self.outputs if self._list_out else self.outputs[0]

    # -- serialisation of utils.save_model / load_model (json + pkl): class names, configs and connectivity
    def get_config(self):
        # Network.get_config: node indices are re-counted over the nodes that belong to THIS network (node_conversion_map)
        mine = {id(n) for n in self._nodes}
        conv = {}
        for layer in self.layers:
            kept = 0
            for i, n in enumerate(layer._inbound_nodes):
                if id(n) in mine:
                    conv[id(n)] = kept
                    kept += 1
        cfg = {'name': self.name, 'layers': []}
        for layer in self.layers:
            nodes = []
            for n in layer._inbound_nodes:
                if id(n) in mine and n.inputs:
                    nodes.append([[s.node.layer.name, conv[id(s.node)], s.idx, {}] for s in n.inputs])
            cfg['layers'].append({'name': layer.name, 'class_name': layer.__class__.__name__, 'config': layer.get_config(),
                                  'inbound_nodes': nodes})
        cfg['input_layers'] = [[s.node.layer.name, 0, 0] for s in self.inputs]
        cfg['output_layers'] = [[s.node.layer.name, conv[id(s.node)], s.idx] for s in self.outputs]
        return cfg

    def to_json(self, **kw):
        import json
        def plain(o):
            if isinstance(o, tuple):
                return list(o)
            if isinstance(o, (np.integer,)):
                return int(o)
            if isinstance(o, (np.floating,)):
                return float(o)
            raise TypeError('Not JSON Serializable: %r' % (o,))
        return json.dumps({'class_name': self.__class__.__name__, 'config': self.get_config(), 'keras_version': '2.2.4',
                           'backend': 'tensorflow'}, default=plain, **kw)


class Model(Network):
    """engine/training.py"""

    def compile(self, optimizer, loss=None, metrics=None, loss_weights=None, **kw):
        from . import losses, metrics as metrics_mod, optimizers
        self.optimizer = optimizers.get(optimizer)
        n_out = len(self.outputs)
        if isinstance(loss, dict):
            loss = [loss.get(n) for n in self.output_names]
        loss_list = list(loss) if isinstance(loss, (list, tuple)) else [loss] * n_out
        self.loss = loss
        self.loss_functions = [losses.get(l) for l in loss_list]
        if isinstance(loss_weights, dict):
            loss_weights = [loss_weights.get(n, 1.) for n in self.output_names]
        self.loss_weights = list(loss_weights) if loss_weights is not None else [1.] * n_out
        self.metrics = metrics or []
        if isinstance(self.metrics, dict):
            per_out = [self.metrics.get(n, []) for n in self.output_names]
            per_out = [m if isinstance(m, (list, tuple)) else [m] for m in per_out]
        else:
            per_out = [list(self.metrics) for _ in range(n_out)]
        self._metric_fns, self.metrics_names = [], ['loss']
        if n_out > 1:
            self.metrics_names += [n + '_loss' for n in self.output_names]
        for i, ms in enumerate(per_out):
            for m in ms:
                fn = metrics_mod.get(m, self.loss_functions[i])
                nm = getattr(fn, '__name__', str(m))
                self.metrics_names.append((self.output_names[i] + '_' + nm) if n_out > 1 else nm)
                self._metric_fns.append((i, fn))
        self._collected = None

    def _feed(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        out = []
        for s, a in zip(self.inputs, xs):
            t = torch.as_tensor(np.asarray(a))
            if t.dim() == len(s._keras_shape) - 1:          # Keras expands (B,) to (B, 1)
                t = t.unsqueeze(-1)
            out.append(t.to(s.example.dtype))
        if len(xs) != len(self.inputs):
            raise ValueError('Error when checking model input: expected %d arrays, got %d' % (len(self.inputs), len(xs)))
        return out

    def _forward(self, x, training):
        outs, _ = self._evaluate(self._feed(x), [None] * len(self.inputs), training)
        return outs

    def predict(self, x, batch_size=32, verbose=0, steps=None):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        n = len(xs[0])
        chunks = []
        with torch.no_grad():
            for i in range(0, max(n, 1), batch_size or n):
                chunks.append([o.cpu().numpy() for o in self._forward([a[i:i + batch_size] for a in xs], False)])
        outs = [np.concatenate([c[j] for c in chunks], 0) for j in range(len(self.outputs))]
        return outs if self._list_out else outs[0]

    predict_on_batch = lambda self, x: self.predict(x, batch_size=len(x[0] if isinstance(x, (list, tuple)) else x))

    def _losses(self, outs, y):
        ys = list(y) if isinstance(y, (list, tuple)) else [y]
        total, parts = 0., []
        for o, t, fn, w in zip(outs, ys, self.loss_functions, self.loss_weights):
            t = torch.as_tensor(np.asarray(t)).to(o.dtype)
            if t.dim() == o.dim() - 1:
                t = t.unsqueeze(-1)
            part = fn(t, o).mean()                  # weighted_masked_objective with no mask / weights: K.mean(score_array)
            parts.append(part)
            total = total + w * part
        mets = []
        with torch.no_grad():
            for i, fn in self._metric_fns:
                t = torch.as_tensor(np.asarray(ys[i])).to(outs[i].dtype)
                mets.append(float(torch.as_tensor(fn(t, outs[i])).to(torch.float64).mean()))
        return total, parts, mets

    def _result(self, total, parts, mets):
        r = [float(total.detach())] + ([float(p.detach()) for p in parts] if len(parts) > 1 else []) + mets
        return r if len(r) > 1 else r[0]

    def unique_trainable_weights(self):
        seen, out = set(), []
        for w in self.trainable_weights:
            if id(w) not in seen:
                seen.add(id(w))
                out.append(w)
        return out

    def train_on_batch(self, x, y, sample_weight=None, class_weight=None):
        params = self.unique_trainable_weights()
        for p in params:
            p.grad = None
        total, parts, mets = self._losses(self._forward(x, True), y)
        grads = torch.autograd.grad(total, params, allow_unused=True)
        grads = [torch.zeros_like(p) if g is None else g for p, g in zip(params, grads)]
        self.last_gradients = {p.vname: g.detach().cpu().numpy().copy() for p, g in zip(params, grads)}
        self.optimizer.apply(params, grads)
        return self._result(total, parts, mets)

    def test_on_batch(self, x, y, sample_weight=None):
        with torch.no_grad():
            total, parts, mets = self._losses(self._forward(x, False), y)
        return self._result(total, parts, mets)

    def evaluate(self, x=None, y=None, batch_size=32, verbose=1, sample_weight=None, steps=None):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        ys = list(y) if isinstance(y, (list, tuple)) else [y]
        n, acc = len(xs[0]), None
        for i in range(0, n, batch_size or n):
            r = self.test_on_batch([np.asarray(a)[i:i + batch_size] for a in xs], [np.asarray(t)[i:i + batch_size] for t in ys])
            r = r if isinstance(r, list) else [r]
            b = min(batch_size or n, n - i)
            acc = [v * b for v in r] if acc is None else [a + v * b for a, v in zip(acc, r)]
        out = [a / n for a in acc]
        return out if len(out) > 1 else out[0]

    def loss_and_gradients(self, x, y, training=True):
        """(shim extension, used by the fixture generator) loss + d loss / d weight by name, no update"""
        params = self.unique_trainable_weights()
        total, _, _ = self._losses(self._forward(x, training), y)
        grads = torch.autograd.grad(total, params, allow_unused=True)
        return float(total.detach()), {p.vname: (np.zeros(tuple(p.shape)) if g is None else g.detach().cpu().numpy().copy())
                              for p, g in zip(params, grads)}

    def fit_generator(self, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
                      validation_steps=None, class_weight=None, max_queue_size=10, workers=1, use_multiprocessing=False,
                      shuffle=True, initial_epoch=0):
        from .callbacks import History
        hist = History()
        hist.model = self
        hist.on_train_begin()
        for cb in callbacks or []:
            cb.model = self
            cb.on_train_begin()
        for epoch in range(initial_epoch, epochs):
            sums = None
            for _ in range(steps_per_epoch):
                x, y = next(generator)[:2]
                r = self.train_on_batch(x, y)
                r = r if isinstance(r, list) else [r]
                sums = r if sums is None else [a + b for a, b in zip(sums, r)]      # Keras logs the running mean
            logs = {k: v / steps_per_epoch for k, v in zip(self.metrics_names, sums)}
            if validation_data is not None:
                ev = self.evaluate_generator(validation_data, validation_steps)
                ev = ev if isinstance(ev, list) else [ev]
                logs.update({'val_' + k: v for k, v in zip(self.metrics_names, ev)})
            hist.on_epoch_end(epoch, logs)
            for cb in callbacks or []:
                cb.on_epoch_end(epoch, logs)
        return hist

    def evaluate_generator(self, generator, steps=None, max_queue_size=10, workers=1, use_multiprocessing=False, verbose=0):
        acc, n = None, 0
        for _ in range(steps):
            x, y = next(generator)[:2]
            r = self.test_on_batch(x, y)
            r = r if isinstance(r, list) else [r]
            b = len(y[0] if isinstance(y, (list, tuple)) else y)
            acc = [v * b for v in r] if acc is None else [a + v * b for a, v in zip(acc, r)]
            n += b
        out = [a / n for a in acc]
        return out if len(out) > 1 else out[0]

    def predict_generator(self, generator, steps=None, **kw):
        outs = [self.predict(next(generator)) for _ in range(steps)]
        if self._list_out:
            return [np.concatenate([o[j] for o in outs], 0) for j in range(len(self.outputs))]
        return np.concatenate(outs, 0)

    def fit(self, x=None, y=None, batch_size=32, epochs=1, verbose=1, callbacks=None, shuffle=True, initial_epoch=0, **kw):
        from .callbacks import History
        hist = History()
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        n = len(xs[0])
        for epoch in range(initial_epoch, epochs):
            order = np.random.permutation(n) if shuffle else np.arange(n)
            tot, cnt = None, 0
            for i in range(0, n, batch_size):
                idx = order[i:i + batch_size]
                yy = [np.asarray(t)[idx] for t in y] if isinstance(y, (list, tuple)) else np.asarray(y)[idx]
                r = self.train_on_batch([np.asarray(a)[idx] for a in xs], yy)
                r = r if isinstance(r, list) else [r]
                tot = [v * len(idx) for v in r] if tot is None else [a + v * len(idx) for a, v in zip(tot, r)]
                cnt += len(idx)
            hist.on_epoch_end(epoch, {k: v / cnt for k, v in zip(self.metrics_names, tot)})
        return hist
