"""keras.models.  TEST INFRASTRUCTURE (oracle/keras_shim/README.md)."""
import json

from . import layers as _layers
from ._engine import Input, InputLayer, Model, Network          # noqa: F401


class Sequential(_layers.Layer):
    """engine/sequential.py, as far as the reference uses it: a stack of layers applied in order, itself usable as a layer
    (task/paper.py:1214-1216 wraps Sequential([Embedding, Reshape]) in TimeDistributed)."""

    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self.layers = []
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        self.layers.append(layer)

    def build(self, input_shape=None):
        self.built = True

    def call(self, inputs, mask=None, training=None):
        x = inputs
        for l in self.layers:
            x = l(x)                        # applied to values: builds on first use
        return x

    def compute_mask(self, inputs, mask=None):
        return None

    @property
    def trainable_weights(self):
        if not self.trainable:
            return []
        return [w for l in self.layers for w in l.trainable_weights]

    @property
    def non_trainable_weights(self):
        out = [w for l in self.layers for w in l.non_trainable_weights]
        if not self.trainable:
            return [w for l in self.layers for w in l.trainable_weights] + out
        return out


def _from_config(config, custom_objects):
    """Network.from_config: rebuild the layers from (class_name, config) and replay the recorded connectivity"""
    scope = dict(vars(_layers))
    scope.update({'Model': Model, 'InputLayer': InputLayer})
    scope.update(custom_objects or {})
    built, tensors = {}, {}

    def make(spec):
        cls, cfg = scope[spec['class_name']], dict(spec['config'])
        if spec['class_name'] == 'Model':
            return _from_config(cfg, custom_objects)
        if spec['class_name'] == 'InputLayer':
            shape = cfg.get('batch_input_shape', [None])[1:]
            return InputLayer(tuple(shape), dtype=cfg.get('dtype'), name=cfg.get('name'))
        if spec['class_name'] == 'TimeDistributed':
            return cls(make(cfg.pop('layer')), name=cfg.get('name'))
        if spec['class_name'] == 'Lambda':
            return cls.from_config(cfg, custom_objects)
        cfg.pop('batch_input_shape', None)
        cfg.pop('dtype', None)
        for k in list(cfg):
            if isinstance(cfg[k], dict) and 'class_name' in cfg[k]:           # serialized initializers / constraints
                cfg.pop(k)
        return cls(**cfg)

    pending = list(config['layers'])
    for spec in pending:
        built[spec['name']] = make(spec)
        if spec['class_name'] == 'InputLayer':
            tensors[(spec['name'], 0, 0)] = built[spec['name']]._inbound_nodes[0].outputs[0]
    progress = True
    done = {}
    while progress:
        progress = False
        for spec in pending:
            layer = built[spec['name']]
            for ni, node in enumerate(spec['inbound_nodes']):
                if (spec['name'], ni) in done:
                    continue
                keys = [(n[0], n[1], n[2]) for n in node]
                if not all(k in tensors for k in keys):
                    continue
                ins = [tensors[k] for k in keys]
                out = layer(ins if len(ins) > 1 else ins[0])
                outs = out if isinstance(out, list) else [out]
                idx = len(layer._inbound_nodes) - 1
                for ti, o in enumerate(outs):
                    tensors[(spec['name'], idx, ti)] = o
                done[(spec['name'], ni)] = True
                progress = True
    ins = [tensors[tuple(k[:3])] for k in config['input_layers']]
    outs = [tensors[tuple(k[:3])] for k in config['output_layers']]
    return Model(ins if len(ins) > 1 else ins[0], outs if len(outs) > 1 else outs[0], name=config.get('name'))


def model_from_json(json_string, custom_objects=None):
    cfg = json.loads(json_string)
    return _from_config(cfg['config'], custom_objects)
