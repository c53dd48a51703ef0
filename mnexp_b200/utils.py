"""Helpers of the reference's utils.py that the LSTUR path uses: Vocab loader and ranking metrics."""
import json
import logging
import pickle

import numpy as np


def load_textual_embedding(path, dimension):
    """Vocab.tsv -> dense (max_index+1, dimension) fp32 matrix, missing rows = 0 (utils.py:34-52)."""
    data = {}
    with open(path) as f:
        for s in f:
            r = s.strip().split('\t')
            if r[-1].count(' ') == dimension - 1:
                data[int(r[-2])] = np.array([float(x) for x in r[-1].split(' ')], dtype=np.float32)
    out = np.zeros((max(data.keys()) + 1, dimension), dtype=np.float32)
    for i, v in data.items():
        out[i] = v
    return out


def dcg_score(y_true, y_score, k=10):
    """utils.py:106-111"""
    order = np.argsort(y_score)[::-1]
    y_true = np.take(y_true, order[:k])
    gains = 2 ** y_true - 1
    discounts = np.log2(np.arange(len(y_true)) + 2)
    return np.sum(gains / discounts)


def ndcg_score(y_true, y_score, k=10):
    """utils.py:114-117"""
    return dcg_score(y_true, y_score, k) / dcg_score(y_true, y_true, k)


def mrr_score(y_true, y_score):
    """utils.py:120-124"""
    order = np.argsort(y_score)[::-1]
    y_true = np.take(y_true, order)
    rr_score = y_true / (np.arange(len(y_true)) + 1)
    return np.sum(rr_score) / np.sum(y_true)


def logging_history(history):
    """utils.py:17-23"""
    for k, v in sorted(history.history.items()):
        logging.info('[*] {}: {}'.format(k, v[-1] if isinstance(v, (list, tuple)) else v))


def logging_evaluation(evaluations):
    """utils.py:26-31"""
    for k, v in sorted(evaluations.items()):
        logging.info('[*] {}: {}'.format(k, v))


# ---- model files: the reference's json + pkl pair (utils.py:66-79, settings model_output) -------------------------
class LoadedModel:
    """What utils.load_model returns here: the architecture record of the json file and the weight list of the pkl.
    `get_weights()` is Keras' call; `apply_to(model)` copies the weights into a built model of the same architecture."""

    def __init__(self, config, weights):
        self.config, self.weights = config, weights

    def get_weights(self):
        return self.weights

    def params(self):
        """{engine parameter name: array} (att_w flattened from Keras' (F, 1) kernel)"""
        out = {}
        for n, a in zip(self.config['weight_names'], self.weights):
            a = np.asarray(a, dtype=np.float32)
            out[n] = a.reshape(-1) if n in ('att_w', 'att_b') else a
        return out

    def apply_to(self, model):
        names = [n for n, _ in model.weight_specs()]
        assert names == self.config['weight_names'], 'architecture mismatch: %s vs %s' % (names, self.config['weight_names'])
        model.set_weights(self.weights)
        return model


def save_model(paths, model):
    """json.dump(model.to_json()) + pickle.dump(model.get_weights(), HIGHEST_PROTOCOL), as utils.py:75-79.  The pkl is
    the same list-of-numpy-arrays a Keras model of this graph pickles; the json string describes the engine's graph
    (class, arch, scorer, weight names and shapes) instead of Keras' layer config."""
    json_path, weight_path = paths
    with open(json_path, 'w') as file:
        json.dump(model.to_json(), file)
    with open(weight_path, 'wb') as file:
        pickle.dump(model.get_weights(), file, protocol=pickle.HIGHEST_PROTOCOL)


# Keras layer class -> the engine's parameter names, in the order keras' get_weights() lists a layer's variables
KERAS_LAYER_WEIGHTS = {
    'Embedding': ('word_emb',),                                   # embeddings (V, E)                  task/paper.py:132-138
    'Conv1D': ('conv_w', 'conv_b'),                               # kernel (k, E, F), bias (F)         task/paper.py:146
    'SimpleAttentionMaskSupport': ('att_w', 'att_b'),             # kernel (F, 1), bias (1)            models.py:446-449
    'Dense': ('dense_w', 'dense_b'),                              # kernel (F, U), bias (U)            task/paper.py:159
}
ENCODER_SHAPES = {'word_emb': 2, 'conv_w': 3, 'conv_b': 1, 'att_w': 2, 'att_b': 1, 'dense_w': 2, 'dense_b': 1}


def _keras_encoder_names(kcfg, weights):
    """Weight names of a Keras-written doc-encoder json (`keras.Model.to_json()` of the model built by
    Seq2VecPaper._get_doc_encoder, task/paper.py:132-160): walk config.layers in order — get_weights() follows the
    layer list — and hand every weighted layer its variables.  Keras itself is not importable here, so the mapping is
    driven by layer class names and checked against the array ranks (unverified against a real Keras
    file); anything unexpected raises instead of guessing."""
    layers = kcfg.get('config', {}).get('layers') if isinstance(kcfg.get('config'), dict) else kcfg.get('config')
    if not isinstance(layers, list):
        raise ValueError('not a Keras model json: no config.layers')
    names = []
    for layer in layers:
        cls = layer.get('class_name')
        if cls in KERAS_LAYER_WEIGHTS:
            if any(n in names for n in KERAS_LAYER_WEIGHTS[cls]):
                raise ValueError('Keras json: a second %s layer — only the cnnatt doc encoder (one Embedding, Conv1D, '
                                 'attention, Dense) can be imported' % cls)
            names += list(KERAS_LAYER_WEIGHTS[cls])
        elif cls in ('GRU', 'LSTM', 'Bidirectional', 'TimeDistributed', 'Model', 'Sequential'):
            raise ValueError('Keras json: layer %s — only the cnnatt doc encoder can be imported' % cls)
    if len(names) != len(weights):
        raise ValueError('Keras json lists %d weighted variables, the pkl holds %d arrays' % (len(names), len(weights)))
    for n, a in zip(names, weights):
        if np.asarray(a).ndim != ENCODER_SHAPES[n]:
            raise ValueError('Keras pkl: %s has rank %d, expected %d' % (n, np.asarray(a).ndim, ENCODER_SHAPES[n]))
    return names


def load_model(paths):
    """Reads a json + pkl pair written by save_model here, or by the reference's utils.save_model for a doc encoder
    (Keras `to_json()` + `get_weights()`; used by --enable-pretrain-encoder, task/paper.py:103-107)."""
    json_path, weight_path = paths
    with open(json_path, 'r') as file:
        config = json.load(file)
        if isinstance(config, str):                 # both writers json.dump a json STRING (utils.py:75-77)
            config = json.loads(config)
    with open(weight_path, 'rb') as file:
        weights = pickle.load(file)
    if 'weight_names' not in config:                # a Keras model json: {"class_name": "Model", "config": {"layers": [...]}}
        config = dict(class_name='keras.' + str(config.get('class_name')), arch='doc_encoder', keras=True,
                      weight_names=_keras_encoder_names(config, weights),
                      weight_shapes=[list(np.asarray(a).shape) for a in weights])
    assert len(weights) == len(config['weight_names'])
    return LoadedModel(config, weights)


# ---- vertical ids of DocMeta column 2 (utils.py:153-155; 'N/A' and unknown names map to 0) ------------------------
VERTICAL_NAMES = ('N/A', 'autos', 'entertainment', 'finance', 'foodanddrink', 'health', 'kids', 'lifestyle', 'movies',
                  'music', 'news', 'sports', 'travel', 'tv', 'video', 'weather')
verticals = {name: i for i, name in enumerate(VERTICAL_NAMES)}


def get_vertical(name):
    return verticals.get(name, 0)
