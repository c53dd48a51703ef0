#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz by running the REFERENCE'S OWN, unmodified source files.

    python tests/golden/make_ref_golden.py [--out FILE]          (needs /root/reference: run in the build container)

`/root/reference/{settings,task/*,models,utils,document}.py` are imported where they lie.  Their two third-party
dependencies, keras 2.2.x and tensorflow 1.x, are neither vendored there nor installable here; `oracle/keras_shim/`
restates the published behaviour of the layers / backend functions those files call (its README.md says exactly what is
restated).  Everything the reference itself wrote is therefore EXECUTED, not re-read: `task.get(config)` loads a synthetic
dataset written in the reference's on-disk formats with its own parsers (document.py, task/seq2vec.py:_load_docs,
task/paper.py:_load_data), `h.train` / `h.valid` are its own window / negative-sampling / pool-shuffle batchers, and
`h.build_model(0)` is its own `_build_model` graph assembly (task/paper.py:222-665).

Per case the file holds: the batch `x, y` the reference's batcher produced (np.random.seed fixed), the weights the model was
given (this repo's parameter names), `model.predict`, `test_model.predict`, the user and candidate vectors, the loss and
d loss / d weight of `model`'s compiled loss in training mode, and the losses + weights of three `train_on_batch` steps with
the reference's compiled `keras.optimizers.Adam`.  Cases with dropout > 0 draw their keep masks from this repo's
counter-based stream (mnexp_b200/rng.py, fp32 verification mode) through the shim's dropout hook, so the CUDA path can be
run on the same masks.

tests/test_ref_pinned.py checks the oracle (CPU) and the CUDA path (-m gpu) against these vectors, and — when
/root/reference is present — re-runs this script in a subprocess and requires the committed file to reproduce.
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('MNEXP_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, 'ref_golden.npz')
GAIN = 1.5
DROP_SEED = 11
VERT_DIM = 3
VSUP_HIDDEN = 10

# (case name, reference task class, reference arch, score model, this repo's oracle arch, extra config)
CASES = [
    ('sid-igru-dot', 'Seq2VecPaperSoftmaxId', 'igru', 'dot', 'igru', {}),
    ('sid-gru-dot', 'Seq2VecPaperSoftmaxId', 'gru', 'dot', 'gru', {}),
    ('sid-ngru-dnn', 'Seq2VecPaperSoftmaxId', 'ngru', 'dnn', 'ngru', {}),
    ('sid-hgru-dot', 'Seq2VecPaperSoftmaxId', 'hgru', 'dot', 'hgru', {}),
    ('sid-dgru-ddot', 'Seq2VecPaperSoftmaxId', 'dgru', 'ddot', 'dgru', {}),
    ('sid-iigru-dot', 'Seq2VecPaperSoftmaxId', 'iigru', 'dot', 'iigru', {}),
    ('sid-vo-dot', 'Seq2VecPaperSoftmaxId', 'vo', 'dot', 'vo', {}),
    ('sid-pgru-dot', 'Seq2VecPaperSoftmaxId', 'pgru', 'dot', 'pgru', {}),
    ('sid-nigru-dot', 'Seq2VecPaperSoftmaxId', 'nigru', 'dot', 'nigru', {}),
    ('sid-niavg-dnn', 'Seq2VecPaperSoftmaxId', 'niavg', 'dnn', 'niavg', {}),
    ('sid-igru-ddot', 'Seq2VecPaperSoftmaxId', 'igru', 'ddot', 'igru', {}),
    ('sid-igru-dot-trainable', 'Seq2VecPaperSoftmaxId', 'igru', 'dot', 'igru', {'textual_embedding_trainable': True}),
    ('sid-igru-dot-dropout', 'Seq2VecPaperSoftmaxId', 'igru', 'dot', 'igru', {'dropout': 0.2}),
    # --enable-pretrain-encoder (task/paper.py:103-107): the doc encoder comes from the json + pkl pair that the reference's
    # own utils.save_model wrote (utils.py:66-79), frozen
    ('sid-igru-dot-pretrain', 'Seq2VecPaperSoftmaxId', 'igru', 'dot', 'igru', {'enable_pretrain_encoder': True}),
    ('s-gru-dot', 'Seq2VecPaperSoftmax', 'gru', 'dot', 'nigru', {}),
    ('s-att-dot', 'Seq2VecPaperSoftmax', 'att', 'dot', 'att', {}),
    ('s-avg-dnn', 'Seq2VecPaperSoftmax', 'avg', 'dnn', 'niavg', {}),
    ('pid-igru', 'Seq2VecPaperId', 'igru', 'dnn', 'igru', {'gain': GAIN}),
    ('pid-iigru', 'Seq2VecPaperId', 'iigru', 'dnn', 'iicat', {'gain': GAIN}),
    ('p-gru', 'Seq2VecPaper', 'gru', 'dnn', 'nigru', {'gain': GAIN}),
    ('pdot-gru', 'Seq2VecPaperDot', 'gru', 'dot', 'nigru', {'gain': GAIN}),
    ('p-att', 'Seq2VecPaper', 'att', 'dnn', 'att', {'gain': GAIN}),
    ('p-avg', 'Seq2VecPaper', 'avg', 'dnn', 'niavg', {'gain': GAIN}),
    ('pid-gru', 'Seq2VecPaperId', 'gru', 'dnn', 'ngru', {'gain': GAIN}),
    ('pid-vo', 'Seq2VecPaperId', 'vo', 'dnn', 'vo', {'gain': GAIN}),
    # users with more than max_impression impressions train on a random.sample of them (task/paper.py:272-279)
    ('pid-igru-maximp', 'Seq2VecPaperId', 'igru', 'dnn', 'igru', {'gain': GAIN, 'max_impression': 1}),
    # time-window batchers (task/paper.py:667-790) and the vertical variants (:793-1001, :1136-1255)
    ('sdays-gru-dot', 'Seq2VecPaperSoftmaxDays', 'gru', 'dot', 'nigru', {'days': 3}),
    ('sdid-igru-dot', 'Seq2VecPaperSoftmaxDaysId', 'igru', 'dot', 'igru', {'days': 3}),
    ('vert-igru-dot', 'Seq2VecPaperSoftmaxDaysIdVert', 'igru', 'dot', 'igru', {'days': 3, 'vertical_embedding_dim': VERT_DIM}),
    ('vert-gru-dnn', 'Seq2VecPaperSoftmaxDaysIdVert', 'gru', 'dnn', 'gru', {'days': 3, 'vertical_embedding_dim': VERT_DIM}),
    ('vsup-igru-dot', 'Seq2VecPaperSoftmaxDaysIdVertSup', 'igru', 'dot', 'igru', {'days': 3, 'hidden_dim': VSUP_HIDDEN, 'gain': 0.5}),
]


def reference_available():
    return os.path.isfile(os.path.join(REF, 'task', 'paper.py'))


def load_reference():
    """put the shim and the reference on sys.path and import the reference's own modules"""
    sys.dont_write_bytecode = True
    sys.path[:0] = [os.path.join(ROOT, 'oracle', 'keras_shim'), REF]
    if ROOT not in sys.path:
        sys.path.append(ROOT)
    import keras
    import tensorflow                      # noqa: F401
    keras.backend.set_floatx('float64')
    import settings
    import task
    return keras, settings, task


def reference_config(settings, data_dir, sh, task_name, arch, score_model, **extra):
    """every option of the reference's CLI (main.py:98-146) at its default, the shape options from `sh`"""
    cfg = dict(task=task_name, arch=arch, round=6, days=30, epochs=2, batch_size=sh.B, training_step=3, validation_step=2,
               validation_impression=5, testing_impression=5, learning_rate=0.001, learning_rate_decay=0.2, gain=1.0,
               window_size=sh.W, dropout=0.0, negative_samples=sh.K, hidden_dim=400, nonlocal_negative_samples=0,
               enable_baseline=False, title_filter_shape=(sh.F, sh.k), title_shape=sh.L, body_shape=sh.L,
               user_embedding_dim=sh.U, textual_embedding_dim=sh.E, textual_embedding_trainable=False, debug=True,
               background=True, name='', pretrain_name='', enable_pretrain_encoder=False, pretrain_encoder_trainable=False,
               personal_embedding_dim=20, news_encoder='cnnatt', score_model=score_model, id_keep=1.0, body_sent_cnt=50,
               body_sent_len=30, body_filter_shape=(400, 3), max_impression=200, max_impression_pos=7, max_impression_neg=200,
               test_window_size=100, vertical_embedding_dim=15, subvertical_embedding_dim=35, use_vertical=False,
               use_vertical_type='vs', use_generator=False, lrd_on_epochs=[1, 3], input_training_data_path=data_dir,
               input_validation_data_path=data_dir, input_previous_model_path=data_dir, output_model_path=data_dir,
               log_dir=data_dir)
    cfg.update(extra)
    with contextlib.redirect_stdout(io.StringIO()):          # Config.__init__ pretty-prints itself
        return settings.Config(cfg)


def _uid(layer):
    return int(layer.name.rsplit('_', 1)[-1]) if layer.name.rsplit('_', 1)[-1].isdigit() else 0


def named_variables(h):
    """this repo's parameter name -> the reference model's weight variable, found by walking the reference's own graph"""
    return named_variables_of(h.model, h)


def named_variables_of(model, h):
    doc = model.get_layer('doc_encoder')
    ue = model.get_layer('user_encoder') if any(l.name == 'user_encoder' for l in model.layers) else None
    out = {}
    for l in doc.layers:
        c = l.__class__.__name__
        if c == 'Embedding':
            out['word_emb'] = l.embeddings
        elif c == 'Conv1D':
            out['conv_w'], out['conv_b'] = l.kernel, l.bias
        elif c == 'SimpleAttentionMaskSupport':
            out['att_w'], out['att_b'] = l.kernel, l.bias
        elif c == 'Dense':
            out['dense_w'], out['dense_b'] = l.kernel, l.bias
    if ue is not None:
        embs = sorted([l for l in ue.layers if l.__class__.__name__ == 'Embedding'], key=_uid)
        for name, l in zip(('user_emb', 'user_emb2'), embs):
            out[name] = l.embeddings
        for l in ue.layers:
            c = l.__class__.__name__
            if c == 'GRU':
                out['gru_wx'], out['gru_wh'], out['gru_b'] = l.kernel, l.recurrent_kernel, l.bias
            elif c == 'Dense':
                out['con_w'], out['con_b'] = l.kernel, l.bias
            elif c == 'SimpleAttentionMaskSupport':
                out['uatt_w'], out['uatt_b'] = l.kernel, l.bias
    for l in model.layers:              # the vertical variants: Sequential([Embedding, Reshape]) / the 'vert' classifier
        inner = getattr(l, 'layer', None)
        if inner is not None and inner.__class__.__name__ == 'Sequential':
            out['vert_emb'] = inner.layers[0].embeddings
        if l.name == 'vert':
            dense = sorted([d for d in inner.layers if d.__class__.__name__ == 'Dense'], key=_uid)
            out['vs_w1'], out['vs_b1'], out['vs_w2'], out['vs_b2'] = dense[0].kernel, dense[0].bias, dense[1].kernel, dense[1].bias
    score_model = getattr(h, 'score_model', None)
    if score_model is not None and not isinstance(score_model, str):          # softmax family: task/paper.py:441-458
        dense = sorted([l for l in score_model.layers if l.__class__.__name__ == 'Dense'], key=_uid)
        names = {'dnn': ('sh', 'so'), 'ddot': ('su', 'sd'), 'dot': ()}[h.config.score_model]
        for n, l in zip(names, dense):
            out[n + '_w'], out[n + '_b'] = l.kernel, l.bias
    else:                                                                      # sigmoid family: task/paper.py:222-226
        for n, lname in (('sh', 'concat_dense'), ('so', 'socre_dense')):
            if any(l.name == lname for l in model.layers):
                l = model.get_layer(lname)
                out[n + '_w'], out[n + '_b'] = l.kernel, l.bias
    return out


def assign(variables, P):
    import keras
    for k, v in variables.items():
        a = np.asarray(P[k], dtype=np.float64)
        if k in ('att_w', 'uatt_w'):
            a = a.reshape(-1, 1)
        assert tuple(v.shape) == a.shape, 'parameter %s: the reference graph has %s, this repo %s' % (k, tuple(v.shape), a.shape)
        keras.backend.set_value(v, a)
    missing = sorted(set(P) - set(variables))
    assert not missing, 'parameters with no variable in the reference graph: %s' % missing


def snapshot(variables):
    return {k: v.detach().cpu().numpy().reshape(-1).copy() if k in ('att_w', 'uatt_w') else v.detach().cpu().numpy().copy()
            for k, v in variables.items()}


class MaskReplay:
    """keep masks of mnexp_b200/rng.py (fp32 verification stream) handed to the shim's Dropout layers in the order the
    reference graph evaluates them: history titles (B*W rows) then candidate k (B rows each); X-dropout has E columns,
    C-dropout F columns (task/paper.py:147,158)."""

    def __init__(self, sh, p=None, seed=None, keep=None):
        from mnexp_b200 import rng
        n = sh.B * (sh.W + 1 + sh.K)
        self.sh, self.calls = sh, {}
        self.keep = keep if keep is not None else {
            sh.E: rng.dropout_multiplier(seed * 2, n * sh.L * sh.E, p).reshape(n, sh.L, sh.E) != 0,
            sh.F: rng.dropout_multiplier(seed * 2 + 1, n * sh.L * sh.F, p).reshape(n, sh.L, sh.F) != 0}

    def reset(self):
        self.calls = {}

    def __call__(self, shape, level):
        sh, width = self.sh, shape[-1]
        keep, nh, c = self.keep[width], sh.B * sh.W, 1 + sh.K
        if shape[0] == nh and (width, 'h') not in self.calls:
            self.calls[(width, 'h')] = 1
            return keep[:nh]
        k = self.calls.get((width, 'c'), 0)
        self.calls[(width, 'c')] = k + 1
        assert shape[0] == sh.B and k < c, (shape, k)
        return keep[nh:].reshape(sh.B, c, sh.L, width)[:, k]


def run_case(mods, sh, data_dir, name, task_name, arch, score_model, my_arch, extra):
    keras, settings, task = mods
    from keras import _engine
    from mnexp_b200 import synth
    keras.backend.clear_session()
    cfg = reference_config(settings, data_dir, sh, task_name, arch, score_model, **extra)
    saved = {}
    if extra.get('enable_pretrain_encoder'):
        import json
        import pickle
        import utils as ref_utils
        cfg0 = reference_config(settings, data_dir, sh, task_name, arch, score_model)
        h0 = task.get(cfg0)
        h0.build_model(0)
        P0 = synth.make_weights(sh, arch=my_arch, bias_noise=0.05, seed=4242, score_model=score_model,
                                word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))
        assign(named_variables(h0), P0)
        ref_utils.save_model(cfg.encoder_input, h0.doc_encoder)          # writes encoder.json / encoder.pkl
        saved['encoder_json'] = np.array(json.load(open(cfg.encoder_input[0])))
        for i, a in enumerate(pickle.load(open(cfg.encoder_input[1], 'rb'))):
            saved['encoder_pkl_%d' % i] = np.asarray(a)
        keras.backend.clear_session()
    h = task.get(cfg)
    model = h.build_model(0)
    variables = named_variables(h)
    for k, v in variables.items():          # layer names may repeat between a loaded encoder and the new layers
        v.vname = k
    softmax = task_name.startswith('Seq2VecPaperSoftmax')
    vert = task_name == 'Seq2VecPaperSoftmaxDaysIdVert'
    vsup = task_name == 'Seq2VecPaperSoftmaxDaysIdVertSup'
    import utils as ref_utils
    kw = {}
    if vert:
        kw['paper_vert'] = cfg.vertical_embedding_dim
    if vsup:
        kw['vertsup'] = (len(ref_utils.verticals), cfg.hidden_dim)
    sm = 'dot' if task_name == 'Seq2VecPaperDot' else score_model
    P = synth.make_weights(sh, arch=my_arch, bias_noise=0.05, seed=4242, score_model=sm,
                           word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')), **kw)
    if 'vert_emb' in P:             # the reference's table has len(utils.verticals) rows
        P['vert_emb'] = P['vert_emb'][:len(ref_utils.verticals)]
    assign(variables, P)
    import random
    np.random.seed(20190131)
    random.seed(20190131)
    gen = h.train
    x, y = next(gen)
    n_cand = 1 + sh.K if softmax else 1
    has_user = any(l.name == 'user_encoder' and len(l.inputs) == 2 for l in model.layers)
    layout = (['user'] if has_user else []) + ['clicked'] + (['clicked_vert'] if vert else []) + ['cand'] * n_cand + \
        (['cand_vert'] * n_cand if vert else [])
    assert len(layout) == len(x), (layout, len(x))
    out = {'x%d' % i: np.asarray(a) for i, a in enumerate(x)}
    out.update(saved)
    out['layout'] = np.array(layout)
    ys = list(y) if isinstance(y, (list, tuple)) else [y]
    for i, a in enumerate(ys):
        out['y' if i == 0 else 'y%d' % i] = np.asarray(a)
    out['n_inputs'], out['n_targets'] = np.int64(len(x)), np.int64(len(ys))
    for k, v in snapshot(variables).items():
        out['P/' + k] = v
    pred = model.predict(x)
    preds = pred if isinstance(pred, list) else [pred]
    for i, a in enumerate(preds):
        out['predict' if i == 0 else 'predict%d' % i] = a
    n_head = layout.index('cand')
    if hasattr(h, 'test_model'):        # [user,] clicked, [clicked_vert,] ONE candidate (the last) [, its vertical]
        one = list(x[:n_head]) + [x[n_head + n_cand - 1]] + ([x[-1]] if vert else [])
        out['test_predict'] = h.test_model.predict(one)
    # intermediate vectors of the outer graph
    doc, has_ue = model.get_layer('doc_encoder'), any(l.name == 'user_encoder' for l in model.layers)
    if has_ue:
        ue = model.get_layer('user_encoder')
        uv = keras.Model(model.inputs[:n_head], ue._inbound_nodes[0].outputs[0])
        out['user_vec'] = uv.predict(list(x[:n_head]))
    out['cand_vec0'] = doc.predict(x[n_head])
    p = float(extra.get('dropout', 0.0))
    replay = MaskReplay(sh, p, DROP_SEED) if p > 0 else None
    _engine._STATE['dropout_hook'] = replay
    try:
        if arch == 'dgru':          # Dropout(0.5, noise_shape=(None, 1)) on the id vector (task/paper.py:608-611)
            loss, grads = model.loss_and_gradients(x, y, training=False)
            keep_rows = (np.random.default_rng(3).random((len(y), 1)) >= 0.5)
            _engine._STATE['dropout_hook'] = lambda shape, level: keep_rows        # the only Dropout with a rate > 0 here
            lt, gt = model.loss_and_gradients(x, y, training=True)
            _engine._STATE['dropout_hook'] = replay
            out['dgru_keep_rows'], out['dgru_train_loss'] = keep_rows.astype(np.float64), np.float64(lt)
            for vname, g in gt.items():
                out['dgru_train_grad/' + vname] = g
        else:
            loss, grads = model.loss_and_gradients(x, y, training=True)
        out['loss'] = np.float64(loss)
        by_var = {v.vname: k for k, v in variables.items()}
        for vname, g in grads.items():
            k = by_var[vname]
            out['grad/' + k] = g.reshape(-1) if k in ('att_w', 'uatt_w') else g
        if arch != 'dgru':
            results = []
            for _ in range(3):
                if replay:
                    replay.reset()
                r = model.train_on_batch(x, y)
                results.append(r if isinstance(r, list) else [r])
            out['adam_losses'] = np.asarray([r[0] for r in results], dtype=np.float64)
            out['adam_results'] = np.asarray(results, dtype=np.float64)           # every entry of metrics_names per step
            out['metrics_names'] = np.array(model.metrics_names)
            for k, v in snapshot(variables).items():
                out['adam/' + k] = v
    finally:
        _engine._STATE['dropout_hook'] = None
    # the NEXT batches of the reference's own batchers (host data path parity: a1-a4 of SURVEY 8)
    x2, y2 = next(gen)
    for i, a in enumerate(x2):
        out['next_x%d' % i] = np.asarray(a)
    for i, a in enumerate(list(y2) if isinstance(y2, (list, tuple)) else [y2]):
        out['next_y' if i == 0 else 'next_y%d' % i] = np.asarray(a)
    np.random.seed(7)
    xv, yv = next(h.valid)
    for i, a in enumerate(xv):
        out['valid_x%d' % i] = np.asarray(a)
    for i, a in enumerate(list(yv) if isinstance(yv, (list, tuple)) else [yv]):
        out['valid_y' if i == 0 else 'valid_y%d' % i] = np.asarray(a)
    # one impression of the reference's own test generator (task/paper.py:415-436 and overrides)
    np.random.seed(9)
    imp = next(iter(h.test_gen()))
    cols = [np.stack(c) for c in zip(*imp)]
    for i, a in enumerate(cols):
        out['test_imp%d' % i] = np.asarray(a)
    out['layers'] = np.array([l.name for l in model.layers])
    out['weight_names'] = np.array([w.vname for w in model.weights])
    return out


# (case name, reference arch, score model, oracle arch, use_vertical_type)  — task/cook.py:4-285
COOK_CASES = [
    ('cook-ingru-ddot-vs', 'ingru', 'ddot', 'igru', 'vs'), ('cook-igru-dnn-v', 'igru', 'dnn', 'ngru', 'v'),
    ('cook-inigru-dot-s', 'inigru', 'ddot', 'iicat', 's'), ('cook-avg-dnn-vs', 'avg', 'dnn', 'niavg', 'vs'),
    ('cook-gru-dot-vs', 'gru', 'dot', 'nigru', 'vs'), ('cook-vo-dnn-vs', 'vo', 'dnn', 'vo', 'vs'),
    ('cook-agru-dot-vs', 'agru', 'dot', 'pgru', 'vs'), ('cook-iavg-dnn-vs', 'iavg', 'dnn', 'iavg', 'vs'),
    ('cook-iatt-ddot-v', 'iatt', 'ddot', 'iatt', 'v'), ('cook-ilstm-dnn-s', 'ilstm', 'dnn', 'ilstm', 's'),
    ('cook-inagru-dot-vs', 'inagru', 'dot', 'inagru', 'vs'), ('cook-atgru-dnn-vs', 'atgru', 'dnn', 'atgru', 'vs'),
    ('cook-algru-dot-vs', 'algru', 'dot', 'algru', 'vs'),
    # id_keep < 1: Dropout(1 - id_keep) on idx_mask, one layer per id table (task/cook.py:141-142, 171-172); training-mode loss
    ('cook-inigru-ddot-s-idkeep', 'inigru', 'ddot', 'iicat', 's'),
]
COOK_ID_KEEP = 0.7
COOK_DV, COOK_DS, COOK_USERS = 3, 5, 25000
MAIN_COOK_BATCH = 10          # 24 training rows -> batches of 10, 10 and a ragged 4 (Keras trains the tail too)


def cook_shape(score_model='dnn', vtype='vs'):
    """'dot' multiplies the user vector with the [title | vert | subvert] news vector directly (task/cook.py:200-201), so
    the reference graph only builds when user_embedding_dim equals that width"""
    from mnexp_b200 import synth
    t = synth.SHAPES['tiny']
    U = t.U if score_model != 'dot' else t.F + (COOK_DV if vtype != 's' else 0) + (COOK_DS if vtype != 'v' else 0)
    return synth.Shape('tinycook', COOK_USERS, t.n_news, t.vocab, L=t.L, W=t.W, K=4, B=t.B, E=t.E, F=t.F, U=U)


def cook_variables(h):
    """this repo's parameter name -> variable, walking the graph Cook._build_model assembled (task/cook.py:214-264)"""
    import utils as ref_utils
    out, seen = {}, set()

    def walk(model, inside_user):
        for l in model.layers:
            inner = getattr(l, 'layer', None)
            if inner is not None and hasattr(inner, 'layers'):
                walk(inner, inside_user)
            if hasattr(l, 'layers') and hasattr(l, 'inputs'):
                is_user = len(l.inputs) == 3 and len(l.inputs[2]._keras_shape) == 3
                walk(l, inside_user or is_user)
            if id(l) in seen:
                continue
            seen.add(id(l))
            c = l.__class__.__name__
            if c == 'Embedding':
                if l.input_dim == COOK_USERS:
                    out.setdefault('_user_embs', []).append(l)
                elif l.input_dim == len(ref_utils.verticals):
                    out['vert_emb'] = l.embeddings
                elif l.input_dim == len(ref_utils.subverticals):
                    out['subvert_emb'] = l.embeddings
                else:
                    out['word_emb'] = l.embeddings
            elif c == 'Conv1D':
                out['conv_w'], out['conv_b'] = l.kernel, l.bias
            elif c == 'SimpleAttentionMaskSupport':
                n = 'uatt' if inside_user else 'att'
                out[n + '_w'], out[n + '_b'] = l.kernel, l.bias
            elif c == 'GRU':
                out['gru_wx'], out['gru_wh'], out['gru_b'] = l.kernel, l.recurrent_kernel, l.bias
            elif c == 'LSTM':
                out['lstm_wx'], out['lstm_wh'], out['lstm_b'] = l.kernel, l.recurrent_kernel, l.bias
            elif c == 'AlphaAdd':
                out['alpha'] = l.alpha
            elif c == 'Dense':
                out.setdefault('_dense', []).append(l)

    walk(h.train_model, False)
    for name, l in zip(('user_emb', 'user_emb2'), sorted(out.pop('_user_embs', []), key=_uid)):
        out[name] = l.embeddings
    dense = sorted(out.pop('_dense', []), key=_uid)
    names = {'dnn': ('sh', 'so'), 'ddot': ('su', 'sd'), 'dot': ()}[h.config.score_model]
    assert len(dense) == len(names), [d.name for d in dense]
    for n, l in zip(names, dense):
        out[n + '_w'], out[n + '_b'] = l.kernel, l.bias
    return out


def run_cook_case(mods, data_dir, name, arch, score_model, my_arch, vtype):
    keras, settings, task = mods
    from mnexp_b200 import synth
    keras.backend.clear_session()
    sh = cook_shape(score_model, vtype)
    id_keep = COOK_ID_KEEP if name.endswith('-idkeep') else 1.0
    cfg = reference_config(settings, data_dir, sh, 'Cook', arch, score_model, use_vertical=True, use_vertical_type=vtype,
                           vertical_embedding_dim=COOK_DV, subvertical_embedding_dim=COOK_DS, days=30, id_keep=id_keep,
                           validation_step=6, lrd_on_epochs=[0])
    h = task.get(cfg)
    model = h.build_model(0)
    variables = cook_variables(h)
    dv = COOK_DV if vtype != 's' else 0
    ds = COOK_DS if vtype != 'v' else 0
    P = synth.make_weights(sh, arch=my_arch, bias_noise=0.05, seed=4242, score_model=score_model, cook=True, dv=dv, ds=ds,
                           word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))
    if not dv:
        P.pop('vert_emb', None)
    if not ds:
        P.pop('subvert_emb', None)
    P['user_emb'] = P['user_emb'] if 'user_emb' in P else None
    P = {k: v for k, v in P.items() if v is not None}
    assign(variables, P)
    x, y = h.train()
    out = {'x%d' % i: np.asarray(a) for i, a in enumerate(x)}
    out['y'] = np.asarray(y[0])
    for k, v in snapshot(variables).items():
        # only the rows of the 25000-row id tables that the data can touch (idx < 50) are stored
        out['P/' + k] = v[:64] if k in ('user_emb', 'user_emb2') else v
    out['predict'] = model.predict(x)
    feats, labels = h.test()
    for i, a in enumerate(feats):
        out['test_x%d' % i] = np.asarray(a)
    out['test_predict'] = h.test_model.predict(feats)
    vx, vy = h.valid()
    out['valid_eval'] = np.asarray(h.test_model.test_on_batch(vx, vy)[:1], dtype=np.float64)     # binary_crossentropy of test_model
    hook_calls = []
    if id_keep < 1.0:       # two Dropout layers on idx_mask ('inigru': one per id table), evaluated in graph order
        from keras import _engine
        g = np.random.default_rng(4)
        keeps = [g.random((len(y[0]), 1)) < id_keep, g.random((len(y[0]), 1)) < id_keep]
        out['idkeep_keep1'], out['idkeep_keep2'] = keeps[0].astype(np.float64), keeps[1].astype(np.float64)

        def hook(shape, level):
            hook_calls.append((shape, level))
            return keeps[(len(hook_calls) - 1) % 2]
        _engine._STATE['dropout_hook'] = hook
    try:
        loss, grads = model.loss_and_gradients(x, y, training=True)
    finally:
        if id_keep < 1.0:
            _engine._STATE['dropout_hook'] = None
    out['loss'] = np.float64(loss)
    out['dropout_calls'] = np.int64(len(hook_calls))
    by_var = {v.vname: k for k, v in variables.items()}
    for vname, g in grads.items():
        k = by_var[vname]
        g = g.reshape(-1) if k in ('att_w', 'uatt_w') else g
        out['grad/' + k] = g[:64] if k in ('user_emb', 'user_emb2') else g
    if id_keep < 1.0:
        out['layers'] = np.array([l.name for l in model.layers])
        return out
    results = [model.train_on_batch(x, y) for _ in range(3)]
    out['adam_results'] = np.asarray(results, dtype=np.float64)
    out['adam_losses'] = out['adam_results'][:, 0]
    for k, v in snapshot(variables).items():
        out['adam/' + k] = v[:64] if k in ('user_emb', 'user_emb2') else v
    lr0 = float(keras.backend.get_value(model.optimizer.lr))
    h.callback(0)                                          # lrd_on_epochs = [0]: learning-rate decay (task/cook.py:279-285)
    out['lr_after_callback'] = np.float64(keras.backend.get_value(model.optimizer.lr))
    out['lr_before_callback'] = np.float64(lr0)
    out['layers'] = np.array([l.name for l in model.layers])
    return out


def run_reference_main(mods, command, cfg, variables_fn, P):
    """Runs the reference's own command function (`main.train` / `main.cook`: its epoch loop, callbacks and evaluation
    tail) and records what it logs.  Instrumentation only: the two logging helpers of utils.py are replaced by recorders,
    and `task.get` hands back the handler with a tap on build_model that loads the seeded weights `P` after the first
    build and fixes numpy's seed — main.py, task/*.py and the rest run unmodified."""
    keras, settings, task = mods
    import logging
    import main as ref_main
    import utils as ref_utils
    records, taps = [], {}
    orig_eval, orig_hist, orig_get = ref_utils.logging_evaluation, ref_utils.logging_history, task.get
    ref_utils.logging_evaluation = lambda d: records.append(('evaluation', {k: float(v) for k, v in d.items()}))
    ref_utils.logging_history = lambda h: records.append(('history', {k: [float(x) for x in v] for k, v in h.history.items()}))

    def tapped_get(config):
        h = orig_get(config)
        build = h.build_model

        def tapped_build(epoch):
            m = build(epoch)
            if epoch == 0:
                taps['variables'] = variables_fn(h)
                assign(taps['variables'], P if P is not None else variables_fn.P)
                np.random.seed(4711)
            return m
        h.build_model = tapped_build
        taps['handler'] = h
        return h
    task.get = tapped_get
    keras.backend.clear_session()
    root = logging.getLogger()
    level = root.level
    np.random.seed(4710)            # handlers may draw at construction (VertAlt shuffles its document split there)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            getattr(ref_main, command).callback(**cfg)
    finally:
        ref_utils.logging_evaluation, ref_utils.logging_history, task.get = orig_eval, orig_hist, orig_get
        for hd in list(root.handlers):
            root.removeHandler(hd)
        root.setLevel(level)
    return taps['handler'], taps['variables'], records


def _flatten_records(records):
    """[(kind, dict)] -> arrays: kinds, keys joined by ',', values concatenated (history lists flattened)"""
    kinds, keys, values, counts = [], [], [], []
    for kind, d in records:
        ks = sorted(d)
        vals = []
        for k in ks:
            v = d[k]
            vals += list(v) if isinstance(v, (list, tuple)) else [v]
        kinds.append(kind)
        keys.append(','.join(ks))
        counts.append(len(vals))
        values += vals
    return dict(kinds=np.array(kinds), keys=np.array(keys), counts=np.array(counts, dtype=np.int64),
                values=np.array(values, dtype=np.float64))


def config_dict(data_dir, sh, task_name, arch, score_model, **extra):
    """the keyword arguments click would hand to the command function: every option of main.py:98-146 + the group's paths"""
    cfg = dict(task=task_name, arch=arch, round=6, days=30, epochs=2, batch_size=sh.B, training_step=3, validation_step=2,
               validation_impression=5, testing_impression=5, learning_rate=0.001, learning_rate_decay=0.2, gain=1.0,
               window_size=sh.W, dropout=0.0, negative_samples=sh.K, hidden_dim=400, nonlocal_negative_samples=0,
               enable_baseline=False, title_filter_shape=(sh.F, sh.k), title_shape=sh.L, body_shape=sh.L,
               user_embedding_dim=sh.U, textual_embedding_dim=sh.E, textual_embedding_trainable=False, debug=True,
               background=True, name='', pretrain_name='', enable_pretrain_encoder=False, pretrain_encoder_trainable=False,
               personal_embedding_dim=20, news_encoder='cnnatt', score_model=score_model, id_keep=1.0, body_sent_cnt=50,
               body_sent_len=30, body_filter_shape=(400, 3), max_impression=200, max_impression_pos=7, max_impression_neg=200,
               test_window_size=100, vertical_embedding_dim=15, subvertical_embedding_dim=35, use_vertical=False,
               use_vertical_type='vs', use_generator=False, lrd_on_epochs=[1, 3], input_training_data_path=data_dir,
               input_validation_data_path=data_dir, input_previous_model_path=data_dir, output_model_path=data_dir,
               log_dir=data_dir)
    cfg.update(extra)
    return cfg


def ref_utils_verticals():
    import utils as ref_utils
    return ref_utils.verticals


def run_main_cases(mods, data_dir, cook_dir):
    """`main.py train` on Seq2VecPaperSoftmaxId (LSTUR-ini) and `main.py cook` on Cook 'ingru', two epochs each"""
    from mnexp_b200 import synth
    keras, settings, task = mods
    out = {}
    sh = synth.SHAPES['tiny']
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=4242, score_model='dot',
                           word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))
    cfg = config_dict(data_dir, sh, 'Seq2VecPaperSoftmaxId', 'igru', 'dot')
    h, variables, records = run_reference_main(mods, 'train', cfg, named_variables, P)
    for k, v in _flatten_records(records).items():
        out['main-train/log_' + k] = v
    for k, v in snapshot(variables).items():
        out['main-train/final/' + k] = v
    for k, v in P.items():
        out['main-train/P/' + k] = np.asarray(v, dtype=np.float64)
    # the sigmoid family inherits Seq2Vec.callback (task/seq2vec.py:296-322) and trains on the weighted BCE
    Pi = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=4242, score_model='dnn',
                            word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))
    cfg = config_dict(data_dir, sh, 'Seq2VecPaperId', 'igru', 'dnn', gain=GAIN)
    h, variables, records = run_reference_main(mods, 'train', cfg, named_variables, Pi)
    for k, v in _flatten_records(records).items():
        out['main-paperid/log_' + k] = v
    for k, v in Pi.items():
        out['main-paperid/P/' + k] = np.asarray(v, dtype=np.float64)
    # the two-output model of ...VertSup through the same loop (five logged metrics per step)
    Ps = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=4242, score_model='dot',
                            vertsup=(len(ref_utils_verticals()), VSUP_HIDDEN), word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))
    cfg = config_dict(data_dir, sh, 'Seq2VecPaperSoftmaxDaysIdVertSup', 'igru', 'dot', days=3, hidden_dim=VSUP_HIDDEN, gain=0.5)
    h, variables, records = run_reference_main(mods, 'train', cfg, named_variables, Ps)
    for k, v in _flatten_records(records).items():
        out['main-vertsup/log_' + k] = v
    for k, v in Ps.items():
        out['main-vertsup/P/' + k] = np.asarray(v, dtype=np.float64)
    # the alternating schedule of ...VertAlt (task/paper.py:1003-1135): round = 3 -> two epochs of the vertical model, one of
    # the click model.  list(set(...)) orders the verticals by string hash, so the classifier columns are stored BY NAME
    import utils as ref_utils
    names = sorted(set(l.split('\t')[2] for l in open(os.path.join(data_dir, 'DocMeta.tsv'))))
    Pv = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=4242, score_model='dot', vertalt=len(names),
                            word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))

    def vertalt_variables(h):
        v = named_variables_of(h.seq_model, h)
        dense = h.vert_model.layers[-1]
        v['vcls_w'], v['vcls_b'] = dense.kernel, dense.bias
        order = [names.index(n) for n in h.verticals]           # reference column j <- vertical h.verticals[j]
        vertalt_variables.P = dict(Pv, vcls_w=Pv['vcls_w'][:, order], vcls_b=Pv['vcls_b'][order])
        return v
    cfg = config_dict(data_dir, sh, 'Seq2VecPaperSoftmaxDaysIdVertAlt', 'igru', 'dot', days=3, round=3, epochs=1)
    h, variables, records = run_reference_main(mods, 'train', cfg, vertalt_variables, None)
    for k, v in _flatten_records(records).items():
        out['main-vertalt/log_' + k] = v
    inv = [h.verticals.index(n) for n in names]                 # back to name order
    for k, v in snapshot(variables).items():
        out['main-vertalt/final/' + k] = v[:, inv] if k == 'vcls_w' else v[inv] if k == 'vcls_b' else v
    for k, v in Pv.items():
        out['main-vertalt/P/' + k] = np.asarray(v, dtype=np.float64)
    out['main-vertalt/vertical_names'] = np.array(names)
    # the vertical-model batchers of ...VertAlt (task/paper.py:1066-1099): document split drawn at construction, train_vert
    # reshuffles it every pass; one-hot columns re-ordered to the sorted vertical names (the reference's order is a set's)
    keras.backend.clear_session()
    np.random.seed(4710)
    hv = task.get(reference_config(settings, data_dir, sh, 'Seq2VecPaperSoftmaxDaysIdVertAlt', 'igru', 'dot', days=3, round=3,
                                   epochs=1))
    hv.build_model(0)
    perm = [hv.verticals.index(n) for n in names]
    np.random.seed(4711)
    gen = hv.train
    for i in range(2):
        tt, vv = next(gen)
        out['main-vertalt/vert_batch%d_titles' % i], out['main-vertalt/vert_batch%d_labels' % i] = np.asarray(tt), np.asarray(vv)[:, perm]
    tt, vv = next(hv.valid)
    out['main-vertalt/vert_valid_titles'], out['main-vertalt/vert_valid_labels'] = np.asarray(tt), np.asarray(vv)[:, perm]
    out['main-vertalt/steps'] = np.array([hv.training_step, hv.validation_step])
    csh = cook_shape('ddot', 'vs')
    Pc = synth.make_weights(csh, arch='igru', bias_noise=0.05, seed=4242, score_model='ddot', cook=True, dv=COOK_DV, ds=COOK_DS,
                            word_emb=np.load(os.path.join(cook_dir, 'Vocab.tsv.npy')))
    cfg = config_dict(cook_dir, csh, 'Cook', 'ingru', 'ddot', use_vertical=True, use_vertical_type='vs', batch_size=MAIN_COOK_BATCH,
                      vertical_embedding_dim=COOK_DV, subvertical_embedding_dim=COOK_DS, validation_step=6, lrd_on_epochs=[0])
    h, variables, records = run_reference_main(mods, 'cook', cfg, cook_variables, Pc)
    for k, v in _flatten_records(records).items():
        out['main-cook/log_' + k] = v
    for k, v in snapshot(variables).items():
        out['main-cook/final/' + k] = v[:64] if k in ('user_emb', 'user_emb2') else v
    for k, v in Pc.items():
        out['main-cook/P/' + k] = np.asarray(v[:64] if k in ('user_emb', 'user_emb2') else v, dtype=np.float64)
    feature, (users, imprs, mask, y_true) = h.test()
    out['main-cook/test_users'], out['main-cook/test_imprs'] = np.asarray(users), np.asarray(imprs)
    out['main-cook/test_mask'], out['main-cook/test_y_true'] = np.asarray(mask), np.asarray(y_true)
    out['main-cook/test_y_pred'] = h.test_model.predict(feature, batch_size=MAIN_COOK_BATCH).reshape(-1)
    return out


# ---- one case at the layer widths of the benchmark (BASELINE.json configs: E300 F400 U200 L30 W50 K4), small tables and
# batch: the tensor-core kernels run their real tile shapes against the reference's own graph.  Weights are re-drawn from
# the same seed by the test (synth.make_weights), big gradients are stored as strided samples.
WIDE_STRIDE = 97


def wide_shape():
    from mnexp_b200 import synth
    return synth.Shape('wide', 300, 400, 2000, L=30, W=50, K=4, B=16, E=300, F=400, U=200)


def wide_weights(sh, word_emb):
    from mnexp_b200 import synth
    return synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=777, score_model='dot', word_emb=word_emb)


def run_wide_case(mods):
    keras, settings, task = mods
    from mnexp_b200 import synth
    keras.backend.clear_session()
    sh = wide_shape()
    data_dir = tempfile.mkdtemp(prefix='refgold_wide_')
    synth.write_dataset(data_dir, sh, seed=17)
    cfg = reference_config(settings, data_dir, sh, 'Seq2VecPaperSoftmaxId', 'igru', 'dot')
    h = task.get(cfg)
    model = h.build_model(0)
    variables = named_variables(h)
    P = wide_weights(sh, np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))
    assign(variables, P)
    np.random.seed(20190131)
    x, y = next(h.train)
    out = {'x%d' % i: np.asarray(a).astype(np.int32) for i, a in enumerate(x)}
    out['y'] = np.asarray(y).astype(np.float32)
    out['predict'] = model.predict(x)
    out['test_predict'] = h.test_model.predict(list(x[:2]) + [x[-1]])
    ue = model.get_layer('user_encoder')
    out['user_vec'] = keras.Model(model.inputs[:2], ue._inbound_nodes[0].outputs[0]).predict(list(x[:2]))
    out['cand_vec0'] = model.get_layer('doc_encoder').predict(x[2])
    loss, grads = model.loss_and_gradients(x, y, training=True)
    out['loss'] = np.float64(loss)
    by_var = {v.vname: k for k, v in variables.items()}
    for vname, g in grads.items():
        k = by_var[vname]
        g = g.reshape(-1)
        out['grad/' + k] = g if g.size <= 4096 else g[::WIDE_STRIDE]
        out['gradnorm/' + k] = np.float64(np.abs(g).max())
    out['adam_losses'] = np.asarray([model.train_on_batch(x, y)[0] for _ in range(3)], dtype=np.float64)
    return out


PIPELINE_CASES = [('pipe-dnn', 'TestPipeline', 'Seq2VecPaper', 'gru', 'dnn', 'nigru'),
                  ('pipe-dot', 'TestPipelineProduct', 'Seq2VecPaperDot', 'gru', 'dot', 'nigru')]


def run_pipeline_case(mods, data_dir, name, pipe_class, task_name, arch, score_model, my_arch):
    """the decomposed scoring pipeline (task/test_pipeline.py): doc vectors once, user vectors from cached doc vectors,
    pair scores, and its own self-check test_correct().  `load_model` (a Keras json + pkl round trip) is replaced by
    handing the pipeline the model object the reference's own _build_model assembled; the stages run unmodified."""
    keras, settings, task = mods
    from mnexp_b200 import synth
    sh = synth.SHAPES['tiny']
    keras.backend.clear_session()
    pipe_dir = tempfile.mkdtemp(prefix='refgold_pipe_')
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab, 8)          # the documents write_dataset(seed=7) wrote
    synth.write_pipeline_files(pipe_dir, tok, sh.W)
    cfg = reference_config(settings, data_dir, sh, task_name, arch, score_model, gain=GAIN, pipeline_input=pipe_dir, name='t')
    h = task.get(cfg)
    h.build_model(0)
    variables = named_variables(h)
    P = synth.make_weights(sh, arch=my_arch, bias_noise=0.05, seed=4242, score_model=score_model,
                           word_emb=np.load(os.path.join(data_dir, 'Vocab.tsv.npy')))
    if score_model == 'dot':
        P = {k: v for k, v in P.items() if not k.startswith(('sh_', 'so_'))}
    assign(variables, P)
    tp = getattr(task, pipe_class)(cfg)
    tp.model, tp.score_encoder = h.model, None
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tp.test_doc_vec()
        tp.test_user_vec()
        tp.test_user_doc_score()
        tp.test_correct()
    out = {'P/' + k: v for k, v in snapshot(variables).items()}
    docs, users = sorted(tp.doc_vec), sorted(tp.user_vec)
    out['doc_keys'], out['doc_vecs'] = np.array(docs), np.stack([tp.doc_vec[d] for d in docs])
    out['user_keys'], out['user_vecs'] = np.array(users), np.stack([tp.user_vec[u] for u in users])
    lines = [l.rstrip('\n').split('\t') for l in open(cfg.pipeline_output)]
    out['score_rows'] = np.array(['\t'.join(l[:3]) for l in lines])
    out['scores'] = np.array([float(l[3]) for l in lines])
    out['stdout'] = np.array(buf.getvalue())
    nums = [float(v) for v in buf.getvalue().replace('[', ' ').replace(']', ' ').split()]
    out['undoc'], out['pred'], out['sigm'] = np.int64(nums[0]), np.float64(nums[1]), np.float64(nums[2])
    return out


HOST_TITLES = ['5 9 2', '7', '1 2 3 4 5 6 7 8 9 10 11 12', '3 4#N#5 6', '12 0 4', '0', '8 8 8 8 8 8 8']
HOST_CONFIG_PROPERTIES = ['training_data_input', 'testing_data_input', 'title_embedding_input', 'doc_meta_input',
                          'user_meta_input', 'model_input', 'model_output', 'encoder_input', 'encoder_output',
                          'user_encoder_output', 'result_output', 'result_input', 'log_output', 'pipeline_inputs',
                          'pipeline_output', 'doc_punc_index_input', 'vertical2idx_input', 'train_npz_input', 'test_npz_input',
                          'train_sparse_input', 'test_sparse_input', 'vert_npz_input']


def run_host_functions(mods, data_dir):
    """the small host-side functions of the path, called as the reference defines them: document parsers (document.py),
    Vocab.tsv loader and ranking metrics (utils.py), vertical tables, the path properties of settings.Config"""
    keras, settings, task = mods
    import document as ref_document
    import utils as ref_utils
    from mnexp_b200 import synth
    sh = synth.SHAPES['tiny']
    out = {}
    parser = ref_document.DocumentParser(ref_document.parse_document(), ref_document.pad_document(1, sh.L))
    out['titles'] = np.array(HOST_TITLES)
    out['parsed_titles'] = np.stack([parser(t)[0] for t in HOST_TITLES])
    out['parsed_dtype'] = np.array(str(parser(HOST_TITLES[0]).dtype))
    out['vocab_tsv'] = ref_utils.load_textual_embedding(os.path.join(data_dir, 'Vocab.tsv'), sh.E)
    g = np.random.default_rng(5)
    scores = [g.random(n) for n in (2, 5, 11, 30)]
    labels = [(g.random(len(s_)) < 0.4).astype(np.float64) for s_ in scores]
    for l_ in labels:
        l_[0] = 1.0
    out['metric_scores'] = np.concatenate(scores)
    out['metric_labels'] = np.concatenate(labels)
    out['metric_lens'] = np.array([len(s_) for s_ in scores])
    out['metric_values'] = np.array([[ref_utils.dcg_score(y, s_, 10), ref_utils.ndcg_score(y, s_, 10), ref_utils.ndcg_score(y, s_, 5),
                                      ref_utils.mrr_score(y, s_)] for s_, y in zip(scores, labels)])
    out['vertical_names'] = np.array(sorted(ref_utils.verticals, key=ref_utils.verticals.get))
    out['subvertical_names'] = np.array(sorted(ref_utils.subverticals, key=ref_utils.subverticals.get))
    out['vertical_lookup'] = np.array([ref_utils.get_vertical(n) for n in ('news', 'sports', 'N/A', 'nope')] +
                                      [ref_utils.get_subvertical(n) for n in ('animals', 'nope')])
    cfg = reference_config(settings, '/data', sh, 'Cook', 'igru', 'dot', days=7, window_size=20, name='n1', pretrain_name='p0',
                           input_previous_model_path='/prev', output_model_path='/out', log_dir='/logs', pipeline_input='/pipe')
    out['config_properties'] = np.array(HOST_CONFIG_PROPERTIES)
    out['config_values'] = np.array([repr(getattr(cfg, p)) for p in HOST_CONFIG_PROPERTIES])
    return out


# (label, task class, arch, score model, extra config): options the reference rejects (SURVEY 8b "Errors")
ERROR_CASES = [('softmaxid-arch', 'Seq2VecPaperSoftmaxId', 'nope', 'dot', {}),
               ('softmax-arch', 'Seq2VecPaperSoftmax', 'igru', 'dot', {}),
               ('paperid-arch', 'Seq2VecPaperId', 'avg', 'dnn', {}),
               ('doc-model', 'Seq2VecPaperSoftmaxId', 'igru', 'dot', {'news_encoder': 'nope'}),
               ('score-model', 'Seq2VecPaperSoftmaxId', 'igru', 'nope', {}),
               ('dot-width', 'Seq2VecPaperSoftmaxId', 'ngru', 'dot', {}),
               ('cook-arch', 'Cook', 'nope', 'dot', {}),
               ('cook-score', 'Cook', 'igru', 'nope', {})]


def run_error_cases(mods, data_dir, cook_dir):
    keras, settings, task = mods
    from mnexp_b200 import synth
    sh = synth.SHAPES['tiny']
    names, kinds, messages = [], [], []
    for label, task_name, arch, score_model, extra in ERROR_CASES:
        keras.backend.clear_session()
        if task_name == 'Cook':
            cfg = reference_config(settings, cook_dir, cook_shape(), 'Cook', arch, score_model, use_vertical=True, days=30,
                                   vertical_embedding_dim=COOK_DV, subvertical_embedding_dim=COOK_DS, **extra)
        else:
            cfg = reference_config(settings, data_dir, sh, task_name, arch, score_model, **extra)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                task.get(cfg).build_model(0)
            kind, msg = 'none', ''
        except Exception as e:          # noqa: BLE001
            kind, msg = type(e).__name__, str(e)
        names.append(label)
        kinds.append(kind)
        messages.append(msg)
    return dict(labels=np.array(names), kinds=np.array(kinds), messages=np.array(messages))


def generate(path=OUT, verbose=True):
    mods = load_reference()
    from mnexp_b200 import synth
    sh = synth.SHAPES['tiny']
    data_dir = tempfile.mkdtemp(prefix='refgold_')
    synth.write_dataset(data_dir, sh)
    out = {'cases': np.array([c[0] for c in CASES]), 'case_table': np.array([[c[0], c[1], c[2], c[3], c[4]] for c in CASES]),
           'gain': np.float64(GAIN), 'drop_seed': np.int64(DROP_SEED)}
    for name, task_name, arch, score_model, my_arch, extra in CASES:
        res = run_case(mods, sh, data_dir, name, task_name, arch, score_model, my_arch, extra)
        for k, v in res.items():
            out[name + '/' + k] = v
        if verbose:
            print('%-24s %-24s loss %.6f  layers %d' % (name, task_name, float(res['loss']), len(res['layers'])))
    cook_dir = tempfile.mkdtemp(prefix='refgold_cook_')
    synth.write_cook_npz(cook_dir, cook_shape())
    out['cook_cases'] = np.array([c[0] for c in COOK_CASES])
    out['cook_table'] = np.array([list(c) for c in COOK_CASES])
    for name, arch, score_model, my_arch, vtype in COOK_CASES:
        res = run_cook_case(mods, cook_dir, name, arch, score_model, my_arch, vtype)
        for k, v in res.items():
            out[name + '/' + k] = v
        if verbose:
            print('%-24s %-24s loss %.6f  layers %d' % (name, 'Cook', float(res['loss']), len(res['layers'])))
    res = run_wide_case(mods)
    for k, v in res.items():
        out['wide/' + k] = v
    if verbose:
        print('%-12s E300 F400 U200 L30 W50 B16: loss %.6f, adam losses %s' % ('wide', float(res['loss']), res['adam_losses']))
    for k, v in run_host_functions(mods, data_dir).items():
        out['host/' + k] = v
    out['pipeline_table'] = np.array([list(c) for c in PIPELINE_CASES])
    for c in PIPELINE_CASES:
        res = run_pipeline_case(mods, data_dir, *c)
        for k, v in res.items():
            out[c[0] + '/' + k] = v
        if verbose:
            print('%-12s %-20s %d docs, %d users, %d scored pairs; test_correct %.8f vs %.8f' % (
                c[0], c[1], len(res['doc_keys']), len(res['user_keys']), len(res['scores']), res['pred'], res['sigm']))
    for k, v in run_error_cases(mods, data_dir, cook_dir).items():
        out['errors/' + k] = v
    if verbose:
        print('errors       ' + ' | '.join('%s: %s' % (a, b) for a, b in zip(out['errors/labels'], out['errors/kinds'])))
    out.update(run_main_cases(mods, data_dir, cook_dir))
    if verbose:
        for c in ('main-train', 'main-paperid', 'main-vertsup', 'main-vertalt', 'main-cook'):
            print('%-12s %d logged records: %s' % (c, len(out[c + '/log_kinds']), ' | '.join(out[c + '/log_keys'][:6])))
    np.savez_compressed(path, **out)
    if verbose:
        print('wrote %d arrays to %s' % (len(out), path))
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=OUT)
    a = ap.parse_args()
    if not reference_available():
        sys.exit('the reference is not present at %s' % REF)
    generate(a.out)
