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


def load_model(paths):
    json_path, weight_path = paths
    with open(json_path, 'r') as file:
        config = json.loads(json.load(file))
    with open(weight_path, 'rb') as file:
        weights = pickle.load(file)
    assert len(weights) == len(config['weight_names'])
    return LoadedModel(config, weights)


# ---- vertical ids of DocMeta column 2 (utils.py:153-155; 'N/A' and unknown names map to 0) ------------------------
VERTICAL_NAMES = ('N/A', 'autos', 'entertainment', 'finance', 'foodanddrink', 'health', 'kids', 'lifestyle', 'movies',
                  'music', 'news', 'sports', 'travel', 'tv', 'video', 'weather')
verticals = {name: i for i, name in enumerate(VERTICAL_NAMES)}


def get_vertical(name):
    return verticals.get(name, 0)
