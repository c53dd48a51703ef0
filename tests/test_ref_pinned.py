"""Parity pinned on the reference's OWN code: tests/golden/ref_golden.npz holds what /root/reference's unmodified
task/paper.py produced (batches of its own batchers, `_build_model` graphs evaluated, losses, gradients, Adam steps) when it
was imported over oracle/keras_shim (tests/golden/make_ref_golden.py explains what is executed and what is restated).

CPU (-m "not gpu"): the oracle (numpy float64 and torch autograd + Keras Adam) and this repo's host data path against
those vectors; when /root/reference is present the generator is re-run in a subprocess and must reproduce the committed
file.  GPU (-m gpu): the CUDA path through the drop-in surface (mnexp_b200.task + keras_like Model protocol) against the
same vectors."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from mnexp_b200 import rng, settings, synth, task
from oracle import lstur_numpy as on
from oracle import lstur_torch as ot

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'ref_golden.npz'))
TABLE = {row[0]: tuple(row[1:]) for row in GOLD['case_table']}
CASES = [str(c) for c in GOLD['cases']]
SH = synth.SHAPES['tiny']
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_ref_golden as mk          # noqa: E402  (numpy-only at import time; the reference is loaded by its functions)
EXTRA = {c[0]: dict(c[5]) for c in mk.CASES}          # the non-default config options of every case
F64 = 1e-9            # float64 restatement against the float64 run of the reference graph
P_KEYS = ('word_emb', 'conv_w', 'conv_b', 'att_w', 'att_b', 'dense_w', 'dense_b', 'user_emb', 'user_emb2', 'gru_wx', 'gru_wh',
          'gru_b', 'con_w', 'con_b', 'uatt_w', 'uatt_b', 'sh_w', 'sh_b', 'so_w', 'so_b', 'su_w', 'su_b', 'sd_w', 'sd_b', 'vert_emb', 'vs_w1', 'vs_b1', 'vs_w2', 'vs_b2')


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def case(name):
    """-> (task class, reference arch, score model, oracle arch, softmax family?, P, user, clicked, cands (B,C,L), y, x)"""
    task_name, arch, score_model, my_arch = TABLE[name]
    g = lambda k: GOLD[name + '/' + k]
    n = int(g('n_inputs'))
    x = [g('x%d' % i) for i in range(n)]
    layout = [str(v) for v in g('layout')]
    softmax = task_name.startswith('Seq2VecPaperSoftmax')
    pick = lambda what: [a for a, l in zip(x, layout) if l == what]
    user = pick('user')[0].astype(np.int64) if 'user' in layout else np.zeros(len(x[0]), dtype=np.int64)
    clicked = pick('clicked')[0].astype(np.int64)
    cands = np.stack(pick('cand'), 1).astype(np.int64)
    P = {k: GOLD[name + '/P/' + k] for k in P_KEYS if name + '/P/' + k in GOLD.files}
    return task_name, arch, score_model, my_arch, softmax, P, user, clicked, cands, g('y'), x


def verticals(name):
    """(hist_vert (B,W), cand_vert (B,C)) integer ids of ...DaysIdVert's extra inputs, or None"""
    g = lambda k: GOLD[name + '/' + k]
    layout = [str(v) for v in g('layout')]
    if 'clicked_vert' not in layout:
        return None, None
    x = [g('x%d' % i) for i in range(len(layout))]
    hv = x[layout.index('clicked_vert')].astype(np.int64)
    cv = np.stack([a.reshape(-1) for a, l in zip(x, layout) if l == 'cand_vert'], 1).astype(np.int64)
    return hv, cv


def vsup_labels(name):
    """integer vertical labels (hist (B,W), cand (B,C)) from the one-hot second target of ...VertSup, or None"""
    if name + '/y1' not in GOLD.files:
        return None
    y1 = GOLD[name + '/y1'].argmax(-1)
    return y1[:, :SH.W], y1[:, SH.W:]


def replayed_masks(p):
    n = SH.B * (SH.W + 1 + SH.K)
    seed = int(GOLD['drop_seed'])
    dx = rng.dropout_multiplier(seed * 2, n * SH.L * SH.E, p).reshape(n, SH.L, SH.E)
    dc = rng.dropout_multiplier(seed * 2 + 1, n * SH.L * SH.F, p).reshape(n, SH.L, SH.F)
    return dx, dc


def oracle_loss(P, name, training_masks=None):
    """the compiled loss of the reference model, restated with the torch oracle; -> (loss, probs, user_vec, cand_vec)"""
    task_name, arch, score_model, my_arch, softmax, _, user, clicked, cands, y, _ = case(name)
    u, c, d = (torch.as_tensor(a).long() for a in (user, clicked, cands))
    B, W, L = c.shape
    C = d.shape[1]
    if training_masks is not None:
        dx, dc = (torch.tensor(m) for m in training_masks)
        nh = B * W
        kw_h = dict(drop_x=dx[:nh], drop_c=dc[:nh])
        kw_c = dict(drop_x=dx[nh:], drop_c=dc[nh:])
    else:
        kw_h = kw_c = {}
    hv, cv = verticals(name)
    dh = ot.news_encoder(c.reshape(B * W, L), P, **kw_h).reshape(B, W, -1)
    dv = ot.news_encoder(d.reshape(B * C, L), P, **kw_c).reshape(B, C, -1)
    if hv is not None:                      # [Dense(U)(title) | Vemb[vertical]], task/paper.py:1218-1232
        dh = torch.cat([dh, P['vert_emb'][torch.as_tensor(hv).long()]], -1)
        dv = torch.cat([dv, P['vert_emb'][torch.as_tensor(cv).long()]], -1)
    H = dh * (c != 0).any(-1).to(dh.dtype).unsqueeze(-1)
    uv = ot.user_encoder(my_arch, u, H, P)
    s = ot.score(uv, dv, P, score_model if (softmax or task_name != 'Seq2VecPaperDot') else 'dot')
    yt = torch.tensor(y, dtype=torch.float64)
    if softmax:
        probs = torch.softmax(s, -1)
        loss = ot.categorical_crossentropy(yt, probs)
        labels = vsup_labels(name)
        if labels is not None:              # loss_weights=[1, gain] over [ranking, vert], task/paper.py:984-990
            vp = ot.vertical_classifier(P, torch.cat([H, dv], 1))
            onehot = torch.tensor(GOLD[name + '/y1'], dtype=torch.float64)
            loss = loss + 0.5 * ot.categorical_crossentropy(onehot, vp)
            return loss, probs, uv, dv, vp
        return loss, probs, uv, dv
    probs = torch.sigmoid(s)
    return ot.weighted_bce(yt.reshape(probs.shape), probs, gain=float(GOLD['gain']), negative_samples=SH.K), probs, uv, dv


@pytest.mark.parametrize('name', CASES)
def test_numpy_oracle_forward_matches_reference_graph(name):
    task_name, arch, score_model, my_arch, softmax, P, user, clicked, cands, y, x = case(name)
    sm = score_model if (softmax or task_name != 'Seq2VecPaperDot') else 'dot'
    hv, cv = verticals(name)
    r = on.lstur_forward(P, user, clicked, cands, arch=my_arch, score_model=sm, aux=True, hist_vert=hv, cand_vert=cv)
    g = lambda k: GOLD[name + '/' + k]
    if name + '/predict1' in GOLD.files:         # second output of ...VertSup: the vertical classifier over [history ; candidates]
        Pt = {k: torch.tensor(v, dtype=torch.float64) for k, v in P.items()}
        vp = oracle_loss(Pt, name)[4]
        assert rel(vp.numpy(), g('predict1')) < F64
    if softmax:
        assert rel(r['probs'], g('predict')) < F64
        # test_model: sigmoid(score_model([user_vec, doc_encoder(candidate)])) on the LAST candidate (task/paper.py:490-495)
        assert rel(r['sigmoid'][:, -1:], g('test_predict')) < F64
    else:
        assert rel(r['sigmoid'], g('predict')) < F64
    if name + '/user_vec' in GOLD.files:
        assert rel(r['user_vec'], g('user_vec')) < F64
    assert rel(r['cand_vec'][:, 0, :g('cand_vec0').shape[1]], g('cand_vec0')) < F64      # doc_encoder output (before any vertical concat)


@pytest.mark.parametrize('name', CASES)
def test_torch_oracle_loss_gradients_and_adam_match_reference_graph(name):
    task_name, arch, score_model, my_arch, softmax, Pn, user, clicked, cands, y, x = case(name)
    g = lambda k: GOLD[name + '/' + k]
    trainable_table = name in EXTRA and EXTRA[name].get('textual_embedding_trainable', False)
    p_drop = EXTRA.get(name, {}).get('dropout', 0.0)
    masks = replayed_masks(p_drop) if p_drop > 0 else None
    frozen = set() if trainable_table else {'word_emb'}
    if EXTRA.get(name, {}).get('enable_pretrain_encoder'):      # encoder.trainable = False, task/paper.py:105-106
        frozen |= {'word_emb', 'conv_w', 'conv_b', 'att_w', 'att_b', 'dense_w', 'dense_b'}
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=k not in frozen) for k, v in Pn.items()}
    trainable = [k for k in P if P[k].requires_grad]
    loss = oracle_loss(P, name, masks)[0]
    assert abs(float(loss.detach()) - float(g('loss'))) < F64
    grads = dict(zip(trainable, torch.autograd.grad(loss, [P[k] for k in trainable], allow_unused=True)))
    checked = 0
    for k in trainable:
        key = name + '/grad/' + k
        assert key in GOLD.files, 'the reference model does not train %s' % k
        got = np.zeros(P[k].shape) if grads[k] is None else grads[k].numpy()
        ref = GOLD[key]
        assert np.abs(got - ref).max() <= F64 * max(1.0, np.abs(ref).max()), k
        checked += 1
    assert checked == len([f for f in GOLD.files if f.startswith(name + '/grad/')])      # same set of trainable tensors
    if arch == 'dgru':
        return
    # three steps of the reference's compiled keras.optimizers.Adam (dense on every tensor, embedding tables included)
    opt = ot.KerasAdam({k: P[k] for k in trainable}, lr=1e-3)
    losses = []
    for _ in range(3):
        loss = oracle_loss(P, name, masks)[0]
        gs = dict(zip(trainable, torch.autograd.grad(loss, [P[k] for k in trainable], allow_unused=True)))
        opt.step(gs)
        losses.append(float(loss))
    assert np.abs(np.array(losses) - g('adam_losses')).max() < F64
    for k in P:
        assert np.abs(P[k].detach().numpy() - g('adam/' + k)).max() < 1e-10, k


# ---------------------------------------------------------------------------------------------- host data path
def _mirror(name, precision='fp32', data_dir=None):
    task_name, arch, score_model, my_arch = TABLE[name]
    d = data_dir or tempfile.mkdtemp()
    synth.write_dataset(d, SH)
    extra = dict(EXTRA.get(name, {}))
    if extra.get('enable_pretrain_encoder'):                     # the json + pkl pair the reference's utils.save_model wrote
        _write_encoder_files(name, d)
        extra.update(input_previous_model_path=d, pretrain_name='')
    cfg = settings.Config(dict(task=task_name, arch=arch, score_model=score_model, input_training_data_path=d,
                               title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B,
                               textual_embedding_dim=SH.E, title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U, debug=True,
                               dropout=extra.pop('dropout', 0.0), precision=precision, validation_impression=5,
                               testing_impression=5, epochs=2, learning_rate=0.001, **extra))
    return task.get(cfg)


def _write_encoder_files(name, d):
    import json
    import pickle
    n = 0
    while '%s/encoder_pkl_%d' % (name, n) in GOLD.files:
        n += 1
    with open(os.path.join(d, 'encoder.json'), 'w') as f:
        json.dump(str(GOLD[name + '/encoder_json']), f)            # utils.save_model json.dumps a json STRING (utils.py:75-77)
    with open(os.path.join(d, 'encoder.pkl'), 'wb') as f:
        pickle.dump([GOLD['%s/encoder_pkl_%d' % (name, i)] for i in range(n)], f, protocol=pickle.HIGHEST_PROTOCOL)
    return os.path.join(d, 'encoder.json'), os.path.join(d, 'encoder.pkl')


def test_model_files_written_by_the_reference_are_importable():
    """utils.save_model of the reference (Keras `to_json()` + `get_weights()` pickle, utils.py:66-79) on its own doc
    encoder -> mnexp_b200.utils.load_model: weight names in Keras' layer order, shapes and values"""
    from mnexp_b200 import utils as mu
    name = 'sid-igru-dot-pretrain'
    loaded = mu.load_model(_write_encoder_files(name, tempfile.mkdtemp()))
    P = loaded.params()
    assert list(P) == ['word_emb', 'conv_w', 'conv_b', 'att_w', 'att_b', 'dense_w', 'dense_b']
    for k, v in P.items():
        ref = GOLD[name + '/P/' + k]
        assert np.array_equal(np.asarray(v, dtype=np.float64).reshape(ref.shape), ref.astype(np.float32).astype(np.float64)) or \
            np.abs(np.asarray(v, dtype=np.float64).reshape(ref.shape) - ref).max() < 1e-7, k


@pytest.mark.parametrize('name', ['sid-igru-dot', 's-gru-dot', 'pid-igru', 'p-gru', 'sdays-gru-dot', 'sdid-igru-dot', 'vert-igru-dot',
                                  'vsup-igru-dot', 'pid-igru-maximp'])
def test_host_batchers_reproduce_the_reference_batches(name):
    """document.py parsers + Window + Impression.negative_samples + the pool-shuffle batcher (`train`) and `valid`: the
    mirror fed the same files and the same numpy seed yields the reference's batches bit for bit."""
    h = _mirror(name)
    g = lambda k: GOLD[name + '/' + k]
    n, nt = int(g('n_inputs')), int(g('n_targets'))
    aslist = lambda y: list(y) if isinstance(y, (list, tuple)) else [y]

    def same(got, prefix, what, count):
        assert len(got) == count, (prefix, what, len(got), count)
        for i, a in enumerate(got):
            ref = g(prefix + what + (str(i) if (what == 'x' or i) else ''))
            assert np.asarray(a).shape == ref.shape and np.array_equal(np.asarray(a), ref), (prefix, what, i)

    import random
    np.random.seed(20190131)
    random.seed(20190131)
    gen = h.train
    for prefix in ('', 'next_'):
        x, y = next(gen)
        same(x, prefix, 'x', n)
        same(aslist(y), prefix, 'y', nt)
    np.random.seed(7)
    xv, yv = next(h.valid)
    same(xv, 'valid_', 'x', n)
    same(aslist(yv), 'valid_', 'y', nt)
    # the first impression of test_gen: rows of ([user,] clicked, [clicked_vert,] title, [vertical,] label)
    np.random.seed(9)
    imp = next(iter(h.test_gen()))
    cols = [np.stack(c) for c in zip(*imp)]
    k = 0
    while name + '/test_imp%d' % k in GOLD.files:
        k += 1
    assert len(cols) == k
    for i, a in enumerate(cols):
        assert np.array_equal(np.asarray(a), g('test_imp%d' % i)), i


def test_committed_vectors_reproduce_from_the_reference():
    """re-run the reference under the shim (subprocess: it shadows `keras` / `tensorflow` / `utils` / `task`)"""
    if not mk.reference_available():
        pytest.skip('the reference tree is not present on this machine')
    out = os.path.join(tempfile.mkdtemp(), 'ref.npz')
    r = subprocess.run([sys.executable, os.path.join(HERE, 'golden', 'make_ref_golden.py'), '--out', out], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    new = np.load(out)
    assert sorted(new.files) == sorted(GOLD.files)
    for k in GOLD.files:
        a, b = GOLD[k], new[k]
        if a.dtype.kind not in 'fc':
            assert np.array_equal(a, b), k
        elif '/adam/' in k or '/final/' in k or k.endswith('log_values'):
            # weights after Adam steps (and what is logged after them): an element whose gradient is at rounding level takes
            # +-lr steps of noise-decided sign (tolerances.assert_adam_weights_close), so a different BLAS thread split may
            # move a few of them; everything else must reproduce to 1e-12
            d = np.abs(a - b)
            assert np.mean(d > 1e-9) <= 1e-3 and (d.size == 0 or d.max() <= 2.1e-2), k
        else:
            assert np.allclose(a, b, rtol=0, atol=1e-12), k


# ---------------------------------------------------------------------------------------------- CUDA path
def _names(model):
    return [k for k in model.WEIGHT_ORDER if k in model._current()]


def _set_weights(model, Pn):
    """the reference model's weights into the mirror's model, by this repo's parameter names"""
    names = _names(model)
    assert sorted(names) == sorted(Pn), (names, sorted(Pn))
    model.set_weights([np.asarray(Pn[k], dtype=np.float32) for k in names])


def _get_weights(model):
    return dict(zip(_names(model), model.get_weights()))


GPU_CASES = [c for c in CASES if c != 'pid-igru-maximp']          # same model as pid-igru: a host-batcher case


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
@pytest.mark.parametrize('name', GPU_CASES)
def test_cuda_path_matches_reference_graph(lib, name, precision):
    """predict / test_model.predict / three train_on_batch steps through the drop-in surface against the vectors of the
    reference's own graph.  fp32 verification mode: 2e-5; tensor-core mode: the 1e-3 of north_star."""
    task_name, arch, score_model, my_arch, softmax, Pn, user, clicked, cands, y, x = case(name)
    if name in EXTRA and 'dropout' in EXTRA[name]:
        pytest.skip('dropout case: test_cuda_dropout_step_matches_reference_graph')
    tol = 2e-5 if precision == 'fp32' else 1e-3
    h = _mirror(name, precision)
    h.config.sparse_user_adam = False                      # the reference's dense Keras-Adam
    model = h.build_model(0)
    _set_weights(model, Pn)
    g = lambda k: GOLD[name + '/' + k]
    same = lambda a, ref: rel(np.asarray(a).reshape(ref.shape), ref)
    layout = [str(v) for v in g('layout')]
    n_head, n_cand = layout.index('cand'), layout.count('cand')
    pred = model.predict(x)
    if name + '/predict1' in GOLD.files:             # [ranking, vert] of ...VertSup
        assert same(pred[0], g('predict')) < tol and same(pred[1], g('predict1')) < tol
        y = [y, g('y1')]
    else:
        assert same(pred, g('predict')) < tol
    if softmax:
        one = list(x[:n_head]) + [x[n_head + n_cand - 1]] + ([x[-1]] if 'cand_vert' in layout else [])
        assert same(h.test_model.predict(one), g('test_predict')) < tol
    if arch == 'dgru':
        return
    results = []
    for _ in range(3):
        r = model.train_on_batch(x, y)
        results.append([float(v) for v in (r if isinstance(r, (list, tuple)) else [r])])
    ref = g('adam_results')
    n_loss = sum(1 for m in g('metrics_names') if str(m).endswith('loss'))
    assert np.abs(np.array(results)[:, :n_loss] - ref[:, :n_loss]).max() < (1e-4 if precision == 'fp32' else 2e-3)
    if softmax and precision == 'fp32':              # categorical accuracies of the same probabilities
        assert np.abs(np.array(results)[:, n_loss:] - ref[:, n_loss:]).max() < 1e-6
    if precision == 'fp32':
        # three dense Keras-Adam steps of lr 1e-3 (|delta| <= 3e-3 per element).  fp32 gradients that agree to 5e-5 keep an
        # element within a small fraction of one step unless its gradient is at rounding level, where Adam's normalisation
        # lets the noise pick the sign (tolerances.assert_adam_weights_close): pooled over the model's elements
        w = _get_weights(model)
        d = np.concatenate([np.abs(np.asarray(w[k], dtype=np.float64).reshape(-1) - g('adam/' + k).reshape(-1)) for k in Pn])
        moved = np.concatenate([np.abs(g('adam/' + k).reshape(-1) - np.asarray(Pn[k], dtype=np.float64).reshape(-1)) for k in Pn])
        assert moved.max() > 1e-3                                  # the reference run did train
        assert np.mean(d > 2e-5) <= 2e-3 and d.max() <= 6.1e-3, (float(d.max()), float(np.mean(d > 2e-5)))


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32'])
def test_cuda_dropout_step_matches_reference_graph(lib, precision):
    """training-mode forward + backward at dropout 0.2 (task/paper.py:147,158): the reference graph was evaluated with the
    keep masks of the device's own counter-based stream (make_ref_golden.MaskReplay), so loss and gradients must agree."""
    from mnexp_b200.engine import LsturEngine
    name = 'sid-igru-dot-dropout'
    task_name, arch, score_model, my_arch, softmax, Pn, user, clicked, cands, y, x = case(name)
    P = {k: np.asarray(v, dtype=np.float32) for k, v in Pn.items()}
    eng = LsturEngine(P, SH.B, SH.W, 1 + SH.K, SH.L, arch=my_arch, dropout=0.2, lr=1e-3, precision=precision,
                      sparse_user_adam=False, score_model=score_model)
    db = eng.to_device_batch(dict(user=user.astype(np.int32), hist_tok=clicked.astype(np.int32), cand_tok=cands.astype(np.int32),
                                  label=np.asarray(y, dtype=np.float32)))
    eng.forward(db, training=True, seed=int(GOLD['drop_seed']))
    eng.backward(db)
    g = lambda k: GOLD[name + '/' + k]
    assert abs(eng.loss() - float(g('loss'))) < 2e-5
    grads = eng.get_grads_dict()
    n = 0
    for k, got in grads.items():
        key = name + '/grad/' + k
        if key in GOLD.files:
            assert rel(got, GOLD[key].reshape(np.asarray(got).shape)) < 5e-5, k
            n += 1
    assert n >= 9


# ---------------------------------------------------------------------------------------------- Cook (task/cook.py:4-285)
COOK = [str(c) for c in GOLD['cook_cases'] if not str(c).endswith('-idkeep')]       # the id_keep case has its own test
COOK_TABLE = {row[0]: tuple(row[1:]) for row in GOLD['cook_table']}
COOK_KEYS = P_KEYS + ('subvert_emb', 'lstm_wx', 'lstm_wh', 'lstm_b', 'alpha')


def cook_case(name):
    arch, score_model, my_arch, vtype = COOK_TABLE[name]
    g = lambda k: GOLD[name + '/' + k]
    x = [g('x%d' % i) for i in range(8)]
    P = {k: GOLD[name + '/P/' + k] for k in COOK_KEYS if name + '/P/' + k in GOLD.files}
    return arch, score_model, my_arch, vtype, P, x, g('y')


def cook_oracle(P, name, x=None, one=False):
    """Cook._build_model restated with the torch oracle -> (logits (n, C), user_vec, cand_vec)"""
    arch, score_model, my_arch, vtype, _, x0, y = cook_case(name)
    idx, idx_mask, ch_title, ch_vert, ch_subvert, cd_title, cd_vert, cd_subvert = x if x is not None else x0
    n = len(idx)
    t = lambda a: torch.as_tensor(np.asarray(a)).long()
    cd_title = np.asarray(cd_title).reshape(n, -1, ch_title.shape[-1])
    kw = dict(arch=my_arch, score_model=score_model, flavour='cook',
              u0_scale=torch.tensor(np.asarray(idx_mask), dtype=torch.float64).reshape(n, 1))
    if 'vert_emb' in P:
        kw.update(hist_vert=np.asarray(ch_vert), cand_vert=np.asarray(cd_vert).reshape(n, -1))
    if 'subvert_emb' in P:
        kw.update(hist_subvert=np.asarray(ch_subvert), cand_subvert=np.asarray(cd_subvert).reshape(n, -1))
    out = ot.forward(P, t(idx).reshape(-1), t(ch_title), t(cd_title), aux=True, **kw)
    return out['logits'], out['user_vec'], out['cand_vec']


@pytest.mark.parametrize('name', COOK)
def test_oracle_matches_reference_cook_graph(name):
    """forward of train_model / test_model, the compiled loss, its gradients and three Adam steps of the reference's own
    Cook._build_model graph (13 user encoders, the three scorers, the three use_vertical_type concatenations)"""
    arch, score_model, my_arch, vtype, Pn, x, y = cook_case(name)
    g = lambda k: GOLD[name + '/' + k]
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=(k != 'word_emb')) for k, v in Pn.items()}
    trainable = [k for k in P if P[k].requires_grad]
    logits, uv, dv = cook_oracle(P, name)
    probs = torch.softmax(logits, -1)
    assert rel(probs.detach().numpy(), g('predict')) < F64
    tx = [g('test_x%d' % i) for i in range(8)]
    tl, _, _ = cook_oracle(P, name, tx)
    assert rel(torch.sigmoid(tl).detach().numpy(), g('test_predict')) < F64
    loss = ot.categorical_crossentropy(torch.tensor(y, dtype=torch.float64), probs)
    assert abs(float(loss.detach()) - float(g('loss'))) < F64
    grads = dict(zip(trainable, torch.autograd.grad(loss, [P[k] for k in trainable], allow_unused=True)))
    n = 0
    for k in trainable:
        key = name + '/grad/' + k
        assert key in GOLD.files, 'the reference model does not train %s' % k
        got = np.zeros(P[k].shape) if grads[k] is None else grads[k].numpy()
        assert np.abs(got - GOLD[key]).max() <= F64 * max(1.0, np.abs(GOLD[key]).max()), k
        n += 1
    assert n == len([f for f in GOLD.files if f.startswith(name + '/grad/')])
    opt = ot.KerasAdam({k: P[k] for k in trainable}, lr=1e-3)
    losses = []
    for _ in range(3):
        logits, _, _ = cook_oracle(P, name)
        loss = ot.categorical_crossentropy(torch.tensor(y, dtype=torch.float64), torch.softmax(logits, -1))
        opt.step(dict(zip(trainable, torch.autograd.grad(loss, [P[k] for k in trainable], allow_unused=True))))
        if 'alpha' in P:                      # models.AlphaAdd: constraint=MinMaxNorm(0, 1), applied after the update
            with torch.no_grad():
                a = P['alpha']
                a.mul_(a.abs().clamp(0.0, 1.0) / (1e-7 + a.abs()))
        losses.append(float(loss.detach()))
    assert np.abs(np.array(losses) - g('adam_losses')).max() < F64
    for k in P:
        assert np.abs(P[k].detach().numpy() - g('adam/' + k)).max() < 1e-10, k
    assert abs(float(g('lr_after_callback')) - float(g('lr_before_callback')) * 0.2) < 1e-15     # lrd_on_epochs, cook.py:279-285


def _cook_mirror(name, precision):
    arch, score_model, my_arch, vtype = COOK_TABLE[name]
    sh = mk.cook_shape(score_model, vtype)
    d = tempfile.mkdtemp()
    synth.write_cook_npz(d, sh)
    cfg = settings.Config(dict(task='Cook', arch=arch, input_training_data_path=d, days=30, window_size=sh.W, batch_size=24,
                               title_filter_shape=(sh.F, 3), user_embedding_dim=sh.U, dropout=0.0, score_model=score_model,
                               use_vertical=True, use_vertical_type=vtype, vertical_embedding_dim=mk.COOK_DV,
                               subvertical_embedding_dim=mk.COOK_DS, precision=precision, validation_step=6,
                               lrd_on_epochs=[0], learning_rate=0.001, id_keep=1.0))
    return task.get(cfg)


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
@pytest.mark.parametrize('name', COOK)
def test_cuda_cook_matches_reference_graph(lib, name, precision):
    """the mirror's Cook handler on the same .npz files and weights: train_model.predict, test_model.predict and three
    full-batch fit() epochs (= three Adam steps) against the reference's own graph"""
    arch, score_model, my_arch, vtype, Pn, x, y = cook_case(name)
    g = lambda k: GOLD[name + '/' + k]
    tol = 2e-5 if precision == 'fp32' else 1e-3
    h = _cook_mirror(name, precision)
    h.config.sparse_user_adam = False
    model = h.build_model(0)
    xm, ym = h.train()
    for a, b in zip(xm, x):                                   # same files -> same features as the reference loaded
        assert np.array_equal(np.asarray(a), b)
    P = {}
    for k, v in h.params.items():
        assert k in Pn, k
        ref = np.asarray(Pn[k], dtype=np.float32)
        if k in ('user_emb', 'user_emb2'):                    # the fixture keeps the 64 rows the data can touch
            full = np.zeros(np.asarray(v).shape, dtype=np.float32)
            full[:len(ref)] = ref
            ref = full
        P[k] = ref.reshape(np.asarray(v).shape)
    assert sorted(P) == sorted(Pn)
    h.params = P
    assert rel(model.predict(xm), g('predict')) < tol
    tx, _ = h.test()
    assert rel(h.test_model.predict(tx).reshape(g('test_predict').shape), g('test_predict')) < tol
    hist = model.fit(xm, ym, 24, epochs=3, initial_epoch=0, shuffle=False)
    assert np.abs(np.array(hist.history['loss']) - g('adam_losses')).max() < (1e-4 if precision == 'fp32' else 2e-3)


# ---------------------------------------------------------------------------------------------- main.py loops
def _records(prefix):
    """[(kind, {key: value | [values]})] as the reference's main.py logged them (make_ref_golden.run_reference_main)"""
    kinds, keys = [str(k) for k in GOLD[prefix + '/log_kinds']], [str(k) for k in GOLD[prefix + '/log_keys']]
    counts, values = GOLD[prefix + '/log_counts'], GOLD[prefix + '/log_values']
    out, pos = [], 0
    for kind, ks, n in zip(kinds, keys, counts):
        ks = ks.split(',')
        vals = values[pos:pos + n]
        pos += n
        per = n // len(ks)
        out.append((kind, {k: vals[i * per:(i + 1) * per] for i, k in enumerate(ks)}))
    return out


def _compare_records(got, ref, loss_tol, rank_tol):
    assert [k for k, _ in got] == [k for k, _ in ref], ([k for k, _ in got], [k for k, _ in ref])
    for (kind, a), (_, b) in zip(got, ref):
        assert sorted(a) == sorted(b), (kind, sorted(a), sorted(b))
        for k in b:
            x, y = np.asarray(a[k], dtype=np.float64).reshape(-1), np.asarray(b[k], dtype=np.float64).reshape(-1)
            if k == 'auc_roc':            # tf.metrics.auc: a streaming 200-bin approximation in the real thing (utils.py:84-96)
                continue
            tol = loss_tol if k.endswith('loss') else rank_tol if k in ('auc', 'mrr', 'ndcgv', 'ndcgx') else 1e-6
            both_nan = np.isnan(x) & np.isnan(y)
            assert np.all(both_nan | (np.abs(x - y) <= tol)), (kind, k, x, y)


def _host_ranking_metrics(scores, labels):
    """(n, 4) auc, ndcg@10, ndcg@5, mrr with the host formulas (sklearn + the utils.py mirrors)"""
    from sklearn.metrics import roc_auc_score
    from mnexp_b200 import utils as mu
    return np.array([[roc_auc_score(y, s), mu.ndcg_score(y, s, 10), mu.ndcg_score(y, s, 5), mu.mrr_score(y, s)]
                     for s, y in zip(scores, labels)])


def test_evaluation_tail_matches_reference_main_cook():
    """main.py:250-297 (per-user / per-impression / in-vocabulary / out-of-vocabulary averages, with the loop's edge
    effects): evaluation.aggregate on the reference run's own predictions reproduces the eight lines it logged."""
    from mnexp_b200 import evaluation
    g = lambda k: GOLD['main-cook/' + k]
    res = evaluation.aggregate(g('test_users'), g('test_imprs'), g('test_mask').reshape(-1), g('test_y_true'), g('test_y_pred'),
                               metric_fn=_host_ranking_metrics)
    got = []
    evaluation.log_aggregate(res, lambda d: got.append(('evaluation', dict(d))))
    ref = _records('main-cook')[-8:]
    _compare_records(got, ref, 1e-9, 1e-6)              # y_pred is stored as float32


def _load_paper_weights(prefix):
    def on_build(h):
        P = {k[len(prefix) + 3:]: GOLD[k] for k in GOLD.files if k.startswith(prefix + '/P/')}
        _set_weights(h.model, P)
        np.random.seed(4711)
    return on_build


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cuda_main_train_loop_matches_reference(lib, precision):
    """`main.py train` on Seq2VecPaperSoftmaxId / igru, two epochs of three steps: fit_generator histories, the callback's
    per-impression AUC / nDCG / MRR (validation and final test impressions), evaluate_generator — every logged number."""
    from mnexp_b200 import main as mmain
    d = tempfile.mkdtemp()
    synth.write_dataset(d, SH)
    cfg = settings.Config(dict(task='Seq2VecPaperSoftmaxId', arch='igru', score_model='dot', input_training_data_path=d,
                               title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B,
                               textual_embedding_dim=SH.E, title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U, debug=True,
                               dropout=0.0, precision=precision, validation_impression=5, testing_impression=5, epochs=2,
                               training_step=3, validation_step=2, learning_rate=0.001, learning_rate_decay=0.2,
                               sparse_user_adam=False))
    np.random.seed(4710)
    h, got = mmain.train(cfg, on_build=_load_paper_weights('main-train'))
    _compare_records(got, _records('main-train'), 1e-4 if precision == 'fp32' else 2e-3, 1e-6 if precision == 'fp32' else 0.2)
    if precision == 'fp32':
        w = _get_weights(h.model if hasattr(h.model, 'WEIGHT_ORDER') else h.test_model)
        d = np.concatenate([np.abs(np.asarray(v, dtype=np.float64).reshape(-1) - GOLD['main-train/final/' + k].reshape(-1))
                            for k, v in w.items()])
        assert np.mean(d > 5e-5) <= 2e-3 and d.max() <= 2.02e-3 * 6, (float(d.max()), float(np.mean(d > 5e-5)))


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cuda_main_cook_loop_matches_reference(lib, precision):
    """`main.py cook` on Cook / ingru / ddot, two epochs of three shuffled batches (10 + 10 + a ragged 4 of the 24 rows): fit histories, test_model.evaluate, the
    learning-rate decay of callback(0), and the eight aggregated evaluation lines of the scored test set."""
    from mnexp_b200 import main as mmain
    sh = mk.cook_shape('ddot', 'vs')
    d = tempfile.mkdtemp()
    synth.write_cook_npz(d, sh)
    cfg = settings.Config(dict(task='Cook', arch='ingru', input_training_data_path=d, days=30, window_size=sh.W,
                               batch_size=mk.MAIN_COOK_BATCH,
                               title_filter_shape=(sh.F, 3), user_embedding_dim=sh.U, dropout=0.0, score_model='ddot',
                               use_vertical=True, use_vertical_type='vs', vertical_embedding_dim=mk.COOK_DV,
                               subvertical_embedding_dim=mk.COOK_DS, precision=precision, validation_step=6, epochs=2,
                               lrd_on_epochs=[0], learning_rate=0.001, learning_rate_decay=0.2, id_keep=1.0,
                               sparse_user_adam=False))

    def on_build(h):
        P = {}
        for k, v in h.params.items():
            ref = np.asarray(GOLD['main-cook/P/' + k], dtype=np.float32)
            if k in ('user_emb', 'user_emb2'):
                full = np.zeros(np.asarray(v).shape, dtype=np.float32)
                full[:len(ref)] = ref
                ref = full
            P[k] = ref.reshape(np.asarray(v).shape)
        h.params = P
        np.random.seed(4711)
    np.random.seed(4710)
    h, got = mmain.cook(cfg, on_build=on_build)
    _compare_records(got, _records('main-cook'), 1e-4 if precision == 'fp32' else 2e-3, 1e-6 if precision == 'fp32' else 0.35)


# ---------------------------------------------------------------------------------------------- decomposed pipeline (C4)
PIPE = {row[0]: tuple(row[1:]) for row in GOLD['pipeline_table']}


def _pipe_inputs():
    """what synth.write_pipeline_files wrote for the reference run: doc tokens, per-user click lists, pairs"""
    d = tempfile.mkdtemp()
    tok, _, _ = synth.make_docs(SH.n_news, SH.L, SH.vocab, 8)
    synth.write_pipeline_files(d, tok, SH.W)
    return d, tok


@pytest.mark.parametrize('name', sorted(PIPE))
def test_oracle_matches_reference_pipeline(name):
    """task/test_pipeline.py run by the reference: doc vectors, user vectors built from CACHED doc vectors (unknown
    documents stay zero, newest click last), raw pair scores — against the oracle on the same files and weights."""
    pipe_class, task_name, arch, score_model, my_arch = PIPE[name]
    g = lambda k: GOLD[name + '/' + k]
    P = {k: GOLD[name + '/P/' + k] for k in P_KEYS if name + '/P/' + k in GOLD.files}
    d, tok = _pipe_inputs()
    doc_ids = {str(k): int(str(k)[1:]) for k in g('doc_keys')}
    dv = on.news_encoder(np.stack([tok[doc_ids[str(k)]] for k in g('doc_keys')]), P)
    assert rel(dv, g('doc_vecs')) < F64
    vec = {str(k): v for k, v in zip(g('doc_keys'), dv)}
    users = {}
    for line in open(os.path.join(d, 'UserClick.tsv')):
        uid, ut, clicks = line.rstrip('\n').split('\t')
        H = np.zeros((SH.W, dv.shape[1]))
        clicks = clicks.split('#N#')
        for i in range(-1, -1 - min(len(clicks), SH.W), -1):
            if clicks[i] in vec:
                H[i] = vec[clicks[i]]
        users[uid + ut] = on.user_encoder(my_arch, np.zeros(1, dtype=int), H[None], P)[0]
    uv = np.stack([users[str(k)] for k in g('user_keys')])
    assert rel(uv, g('user_vecs')) < F64
    for row, sc in zip(g('score_rows'), g('scores')):
        uid, ut, doc = str(row).split('\t')
        s = on.score(users[uid + ut][None], vec[doc][None, None], P, score_model)[0, 0]
        assert abs(s - sc) < 1e-9
    assert abs(float(g('pred')) - float(g('sigm'))) < 1e-7 and int(g('undoc')) == 1          # the reference's own self-check


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
@pytest.mark.parametrize('name', sorted(PIPE))
def test_cuda_pipeline_matches_reference(lib, name, precision):
    from mnexp_b200.task.test_pipeline import TestPipeline, TestPipelineProduct
    pipe_class, task_name, arch, score_model, my_arch = PIPE[name]
    g = lambda k: GOLD[name + '/' + k]
    tol = 2e-5 if precision == 'fp32' else 1e-3
    pd, _ = _pipe_inputs()
    dd = tempfile.mkdtemp()
    synth.write_dataset(dd, SH)
    cfg = settings.Config(dict(task=task_name, arch=arch, score_model=score_model, input_training_data_path=dd,
                               title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B,
                               textual_embedding_dim=SH.E, title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U, debug=True,
                               dropout=0.0, precision=precision, gain=float(GOLD['gain']), pipeline_input=pd, name='t'))
    h = task.get(cfg)
    model = h.build_model(0)
    _set_weights(model, {k: GOLD[name + '/P/' + k] for k in P_KEYS if name + '/P/' + k in GOLD.files})
    tp = (TestPipelineProduct if pipe_class == 'TestPipelineProduct' else TestPipeline)(cfg)
    tp.load_model(h)
    tp.test_doc_vec()
    tp.test_user_vec()
    tp.test_user_doc_score()
    assert sorted(tp.doc_vec) == sorted(str(k) for k in g('doc_keys')) and sorted(tp.user_vec) == sorted(str(k) for k in g('user_keys'))
    assert rel(np.stack([tp.doc_vec[str(k)] for k in g('doc_keys')]), g('doc_vecs')) < tol
    assert rel(np.stack([tp.user_vec[str(k)] for k in g('user_keys')]), g('user_vecs')) < tol
    lines = [l.rstrip('\n').split('\t') for l in open(cfg.pipeline_output)]
    assert ['\t'.join(l[:3]) for l in lines] == [str(r) for r in g('score_rows')]
    got = np.array([float(l[3]) for l in lines])
    assert np.abs(got - g('scores')).max() <= tol * max(1.0, np.abs(g('scores')).max())
    pred, sigm = tp.test_correct()
    assert abs(float(np.asarray(pred).reshape(-1)[0]) - float(g('pred'))) < tol
    assert abs(float(np.asarray(sigm).reshape(-1)[0]) - float(g('sigm'))) < tol


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cuda_main_vertalt_schedule_matches_reference(lib, precision):
    """`main.py train` on Seq2VecPaperSoftmaxDaysIdVertAlt with round = 3 (task/paper.py:1003-1135): two epochs of the
    vertical classifier on the shared doc encoder (its own document split, steps and learning-rate decay), then one epoch
    of the click model with the callback's evaluation — the reference's whole log."""
    from mnexp_b200 import main as mmain
    d = tempfile.mkdtemp()
    synth.write_dataset(d, SH)
    cfg = settings.Config(dict(task='Seq2VecPaperSoftmaxDaysIdVertAlt', arch='igru', score_model='dot', input_training_data_path=d,
                               title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B, days=3, round=3,
                               textual_embedding_dim=SH.E, title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U, debug=True,
                               dropout=0.0, precision=precision, validation_impression=5, testing_impression=5, epochs=1,
                               training_step=3, validation_step=2, learning_rate=0.001, learning_rate_decay=0.2,
                               sparse_user_adam=False))

    def on_build(h):
        assert list(h.verticals) == [str(n) for n in GOLD['main-vertalt/vertical_names']]
        P = {k[len('main-vertalt/P/'):]: GOLD[k] for k in GOLD.files if k.startswith('main-vertalt/P/')}
        _set_weights(h.seq_model, P)
        np.random.seed(4711)
    np.random.seed(4710)
    h, got = mmain.train(cfg, on_build=on_build)
    _compare_records(got, _records('main-vertalt'), 1e-4 if precision == 'fp32' else 2e-3, 1e-6 if precision == 'fp32' else 0.35)


# ---------------------------------------------------------------------------------------------- benchmark layer widths
def _wide():
    sh = mk.wide_shape()
    P = mk.wide_weights(sh, synth.make_vocab(sh.vocab, sh.E, 17))
    g = lambda k: GOLD['wide/' + k]
    x = [g('x%d' % i) for i in range(7)]
    return sh, P, g, x[0].astype(np.int64), x[1].astype(np.int64), np.stack(x[2:], 1).astype(np.int64), g('y')


def _strided(a):
    a = np.asarray(a).reshape(-1)
    return a if a.size <= 4096 else a[::mk.WIDE_STRIDE]


def test_oracle_matches_reference_graph_at_benchmark_widths():
    """E300 F400 U200 L30 W50 K4 (the layer widths of BASELINE.json's configs), B = 16: the reference's own graph vs the
    torch oracle — forward, loss, every gradient (large ones on a stride-97 sample), three Adam steps."""
    sh, Pn, g, user, clicked, cands, y = _wide()
    ora = ot.LsturOracle(Pn, arch='igru', lr=1e-3)
    out = ora.forward(user, clicked, cands, aux=True)
    assert rel(out['probs'].detach().numpy(), g('predict')) < F64
    assert rel(out['user_vec'].detach().numpy(), g('user_vec')) < F64
    assert rel(out['cand_vec'][:, 0].detach().numpy(), g('cand_vec0')) < F64
    assert rel(torch.sigmoid(out['logits'][:, -1:]).detach().numpy(), g('test_predict')) < F64
    loss, grads = ora.loss_and_grads(user, clicked, cands)
    assert abs(float(loss) - float(g('loss'))) < F64
    for k, gr in grads.items():
        ref = g('grad/' + k)
        assert np.abs(_strided(gr.numpy()) - ref).max() <= 1e-9 * max(1.0, float(g('gradnorm/' + k))), k
    losses = [ora.train_step(user, clicked, cands, training=False) for _ in range(3)]
    assert np.abs(np.array(losses) - g('adam_losses')).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cuda_matches_reference_graph_at_benchmark_widths(lib, precision):
    """the same case through the C ABI: the tcgen05 conv / weight-gradient / GRU kernels at their real tile shapes against
    the reference's own graph (fp16_tc: 1e-3 on outputs and loss, gradients to the 16-bit operand rounding)"""
    from mnexp_b200.engine import LsturEngine
    from tolerances import assert_close
    sh, Pn, g, user, clicked, cands, y = _wide()
    tc = precision != 'fp32'
    tol = 1e-3 if tc else 2e-5
    eng = LsturEngine(Pn, sh.B, sh.W, 1 + sh.K, sh.L, arch='igru', lr=1e-3, precision=precision, sparse_user_adam=False)
    db = eng.to_device_batch(dict(user=user.astype(np.int32), hist_tok=clicked.astype(np.int32), cand_tok=cands.astype(np.int32),
                                  label=np.asarray(y, dtype=np.float32)))
    probs = eng.forward(db, training=True).cpu().numpy().copy()
    assert_close(probs, g('predict'), tol, 'probs')
    assert_close(eng.view('user_vec').reshape(sh.B, -1).cpu().numpy(), g('user_vec'), tol, 'user_vec')
    assert_close(eng.score_sigmoid().reshape(sh.B, -1)[:, -1:].cpu().numpy(), g('test_predict'), tol, 'test_model score')
    assert abs(eng.loss() - float(g('loss'))) < tol
    eng.backward(db)
    got = eng.get_grads_dict()
    n = 0
    for k, v in got.items():
        if 'wide/grad/' + k not in GOLD.files:
            continue
        ref, scale = g('grad/' + k), float(g('gradnorm/' + k))
        err = np.abs(_strided(v) - ref).max() / max(scale, 1e-30)
        # att_b: the sum of d a over every token cancels almost exactly (softmax constraint): scale of its sibling att_w
        if k == 'att_b':
            err = np.abs(_strided(v) - ref).max() / float(g('gradnorm/att_w'))
        # tensor-core mode: 16-bit operands flip a ~1e-3 fraction of near-zero ReLU gates relative to float64, which moves the
        # title-encoder gradients by a few % of their max-norm (proven there by shifting the bias); the rest 3e-2
        gtol = 5e-5 if not tc else (6e-2 if k in ('conv_w', 'conv_b', 'att_w', 'att_b') else 3e-2)
        assert err < gtol, (k, err)
        n += 1
    assert n >= 9
    losses = [float(eng.train_step(db)[0]) for _ in range(3)]
    assert np.abs(np.array(losses) - g('adam_losses')).max() < (1e-4 if not tc else 3e-3)


def _paper_cfg(task_name, arch, score_model, precision, **extra):
    d = tempfile.mkdtemp()
    synth.write_dataset(d, SH)
    return settings.Config(dict(task=task_name, arch=arch, score_model=score_model, input_training_data_path=d,
                                title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B,
                                textual_embedding_dim=SH.E, title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U, debug=True,
                                dropout=0.0, precision=precision, validation_impression=5, testing_impression=5, epochs=2,
                                training_step=3, validation_step=2, learning_rate=0.001, learning_rate_decay=0.2,
                                sparse_user_adam=False, **extra))


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cuda_main_train_loop_sigmoid_family_matches_reference(lib, precision):
    """`main.py train` on Seq2VecPaperId (weighted BCE, Seq2Vec.callback of task/seq2vec.py:296-322: learning-rate decay +
    ranking metrics over `testing_impression` impressions)"""
    from mnexp_b200 import main as mmain
    cfg = _paper_cfg('Seq2VecPaperId', 'igru', 'dnn', precision, gain=float(GOLD['gain']))
    np.random.seed(4710)
    h, got = mmain.train(cfg, on_build=_load_paper_weights('main-paperid'))
    _compare_records(got, _records('main-paperid'), 1e-4 if precision == 'fp32' else 2e-3, 1e-6 if precision == 'fp32' else 0.35)


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cuda_main_train_loop_vertsup_matches_reference(lib, precision):
    """`main.py train` on ...DaysIdVertSup: the two-output model through fit_generator / evaluate_generator (five logged
    metrics per step: loss, ranking_loss, vert_loss and the two accuracies) and the callback's evaluation"""
    from mnexp_b200 import main as mmain
    cfg = _paper_cfg('Seq2VecPaperSoftmaxDaysIdVertSup', 'igru', 'dot', precision, days=3, hidden_dim=mk.VSUP_HIDDEN, gain=0.5)
    np.random.seed(4710)
    h, got = mmain.train(cfg, on_build=_load_paper_weights('main-vertsup'))
    _compare_records(got, _records('main-vertsup'), 1e-4 if precision == 'fp32' else 3e-3, 1e-6 if precision == 'fp32' else 0.35)


# ---------------------------------------------------------------------------------------------- small host functions
def test_host_functions_match_the_reference():
    """document.py parsers (float64 rows, right zero padding, truncation, first non-empty clause only), the Vocab.tsv loader,
    utils.dcg / ndcg / mrr, the vertical tables and every path property of settings.Config — outputs of the reference's own
    functions (make_ref_golden.run_host_functions) against the mirror's"""
    from mnexp_b200 import document, utils as mu
    g = lambda k: GOLD['host/' + k]
    parser = document.DocumentParser(document.parse_document(), document.pad_document(1, SH.L))
    got = np.stack([parser(str(t))[0] for t in g('titles')])
    assert str(got.dtype) == str(g('parsed_dtype')) and np.array_equal(got, g('parsed_titles'))
    d = tempfile.mkdtemp()
    synth.write_dataset(d, SH)
    emb = mu.load_textual_embedding(os.path.join(d, 'Vocab.tsv'), SH.E)
    assert emb.dtype == g('vocab_tsv').dtype and np.array_equal(emb, g('vocab_tsv'))
    pos = 0
    for n, ref in zip(g('metric_lens'), g('metric_values')):
        s, y = g('metric_scores')[pos:pos + n], g('metric_labels')[pos:pos + n]
        pos += n
        mine = [mu.dcg_score(y, s, 10), mu.ndcg_score(y, s, 10), mu.ndcg_score(y, s, 5), mu.mrr_score(y, s)]
        assert np.abs(np.array(mine) - ref).max() < 1e-15
    # (the 306-name subvertical table, utils.py:157-228, is not mirrored: the paper classes never read News.subvertical and
    # Cook receives integer ids in its .npz files; only its SIZE, 307 rows, is part of the model — checked by the Cook cases)
    look = [mu.get_vertical(n) for n in ('news', 'sports', 'N/A', 'nope')]
    assert look == [int(v) for v in g('vertical_lookup')[:4]]
    assert [str(n) for n in g('vertical_names')] == sorted(mu.verticals, key=mu.verticals.get)
    assert len(g('subvertical_names')) == 307
    cfg = settings.Config(dict(task='Cook', arch='igru', input_training_data_path='/data', days=7, window_size=20, name='n1',
                               pretrain_name='p0', input_previous_model_path='/prev', output_model_path='/out', log_dir='/logs',
                               pipeline_input='/pipe'))
    out_of_scope = {'user_meta_input', 'user_encoder_output', 'result_output', 'result_input', 'doc_punc_index_input',
                    'vertical2idx_input', 'train_sparse_input', 'test_sparse_input', 'vert_npz_input'}     # other experiments' files
    for prop, ref in zip(g('config_properties'), g('config_values')):
        if hasattr(type(cfg), str(prop)):
            assert repr(getattr(cfg, str(prop))) == str(ref), prop
        else:
            assert str(prop) in out_of_scope, prop


def test_training_parity_fixture_matches_reference_graph():
    """north_star: "AUC within 0.002 after a fixed step count".  tests/golden/train_parity_c1.npz — what the CUDA arms are
    compared with in tests/test_gpu_training_parity.py — was trained by the oracle; train_parity_c1_ref.npz is the same run
    (C1, 200 steps, same batches and initial weights, float32) trained by the REFERENCE'S OWN graph and compiled Adam over
    the shim (tests/golden/make_train_parity_ref.py).  The two agree: per-step losses (fp32 reassociation grows over the
    200 steps) and the held-out per-impression AUC."""
    a = np.load(os.path.join(HERE, 'golden', 'train_parity_c1.npz'))
    b = np.load(os.path.join(HERE, 'golden', 'train_parity_c1_ref.npz'))
    # p0: dropout 0; p2: dropout 0.2, the device's tensor-core keep masks of every step fed to the reference graph's own
    # Dropout layers through the shim's hook (measured: losses within 8e-4 / 1.2e-3 over the 200 steps, AUC 0.866455 vs
    # 0.866455 and 0.86560 vs 0.86511)
    for run in ('p0', 'p2'):
        d = np.abs(a['loss_' + run] - b['ref_loss_' + run])
        assert len(d) == 200 and d[:10].max() < 1e-4 and d.max() < 5e-3, run
        assert abs(float(a['auc_' + run]) - float(b['ref_auc_' + run])) <= 0.002, run
        assert abs(synth.impression_auc(b['ref_probs_' + run]) - float(b['ref_auc_' + run])) < 1e-12
        assert float(b['ref_auc_' + run]) > float(a['auc_init']) + 0.1          # the task was learnt



# ---------------------------------------------------------------------------------------------- dropout on the id vector
def test_oracle_matches_reference_id_vector_dropout():
    """Training-mode dropout of the user-id vector, keep draws fed to the reference graph through the shim's hook:
    `dgru` — Dropout(0.5, noise_shape=(None, 1)) on the id vector, one draw per row (task/paper.py:608-611);
    Cook `inigru` with id_keep = 0.7 — Dropout(0.3) on idx_mask, one INDEPENDENT layer per id table (task/cook.py:141-142,
    171-172).  The oracle takes the same draws as explicit multipliers (u0_scale / u2_scale = what the engine's
    lstur_batch.user_scale / user_scale2 carry)."""
    # dgru
    name = 'sid-dgru-ddot'
    task_name, arch, score_model, my_arch, softmax, Pn, user, clicked, cands, y, x = case(name)
    g = lambda k: GOLD[name + '/' + k]
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=(k != 'word_emb')) for k, v in Pn.items()}
    scale = torch.tensor(g('dgru_keep_rows') / 0.5)
    u, c, d = (torch.as_tensor(a).long() for a in (user, clicked, cands))
    probs = ot.forward(P, u, c, d, arch=my_arch, score_model=score_model, u0_scale=scale)
    loss = ot.categorical_crossentropy(torch.tensor(y, dtype=torch.float64), probs)
    assert abs(float(loss.detach()) - float(g('dgru_train_loss'))) < F64
    assert abs(float(g('dgru_train_loss')) - float(g('loss'))) > 1e-6             # the draw did change the loss
    tr = [k for k in P if P[k].requires_grad]
    for k, gr in zip(tr, torch.autograd.grad(loss, [P[k] for k in tr], allow_unused=True)):
        ref = GOLD['%s/dgru_train_grad/%s' % (name, k)]
        got = np.zeros(ref.shape) if gr is None else gr.numpy().reshape(ref.shape)
        assert np.abs(got - ref).max() <= F64 * max(1.0, np.abs(ref).max()), k
    # cook inigru, id_keep 0.7
    name = 'cook-inigru-ddot-s-idkeep'
    g = lambda k: GOLD[name + '/' + k]
    assert int(g('dropout_calls')) == 2                                          # two Dropout layers, two draws
    arch, score_model, my_arch, vtype = [str(v) for v in GOLD['cook_table'][[str(c) for c in GOLD['cook_cases']].index(name)][1:]]
    xs = [g('x%d' % i) for i in range(8)]
    Pn = {k: GOLD[name + '/P/' + k] for k in COOK_KEYS if name + '/P/' + k in GOLD.files}
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=(k != 'word_emb')) for k, v in Pn.items()}
    idx, idx_mask, ch_title, ch_vert, ch_subvert, cd_title, cd_vert, cd_subvert = xs
    n = len(idx)
    keep = mk.COOK_ID_KEEP
    s1 = torch.tensor(idx_mask.reshape(n, 1) * g('idkeep_keep1') / keep)
    s2 = torch.tensor(idx_mask.reshape(n, 1) * g('idkeep_keep2') / keep)
    t = lambda a: torch.as_tensor(np.asarray(a)).long()
    out = ot.forward(P, t(idx).reshape(-1), t(ch_title), t(cd_title), arch=my_arch, score_model=score_model, flavour='cook',
                     aux=True, u0_scale=s1, u2_scale=s2, hist_subvert=ch_subvert, cand_subvert=cd_subvert)
    loss = ot.categorical_crossentropy(torch.tensor(g('y'), dtype=torch.float64), torch.softmax(out['logits'], -1))
    assert abs(float(loss.detach()) - float(g('loss'))) < F64
    tr = [k for k in P if P[k].requires_grad]
    for k, gr in zip(tr, torch.autograd.grad(loss, [P[k] for k in tr], allow_unused=True)):
        ref = g('grad/' + k)
        got = np.zeros(P[k].shape) if gr is None else gr.numpy()
        assert np.abs(got.reshape(ref.shape) - ref).max() <= F64 * max(1.0, np.abs(ref).max()), k



def test_vertalt_vertical_batchers_reproduce_the_reference():
    """...VertAlt's own data path (task/paper.py:1009-1099): the 10 % / 90 % document split shuffled at construction, the
    reshuffling `train_vert`, `valid_vert`, and the step counts derived from them — same numpy seed, same batches (one-hot
    columns compared in the order of the sorted vertical names; the reference orders them by a set's iteration)"""
    d = tempfile.mkdtemp()
    synth.write_dataset(d, SH)
    cfg = settings.Config(dict(task='Seq2VecPaperSoftmaxDaysIdVertAlt', arch='igru', score_model='dot', input_training_data_path=d,
                               title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B, days=3, round=3,
                               textual_embedding_dim=SH.E, title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U, debug=True,
                               dropout=0.0, epochs=1, training_step=3, validation_step=2))
    np.random.seed(4710)
    h = task.get(cfg)
    g = lambda k: GOLD['main-vertalt/' + k]
    assert list(h.verticals) == [str(n) for n in g('vertical_names')]
    assert [h.training_step, h.validation_step] == [int(v) for v in g('steps')]
    np.random.seed(4711)
    gen = h.train                      # train_seq is False before the first callback_valid: vertical batches
    for i in range(2):
        tt, vv = next(gen)
        assert np.array_equal(np.asarray(tt), g('vert_batch%d_titles' % i)) and np.array_equal(np.asarray(vv), g('vert_batch%d_labels' % i))
    tt, vv = next(h.valid)
    assert np.array_equal(np.asarray(tt), g('vert_valid_titles')) and np.array_equal(np.asarray(vv), g('vert_valid_labels'))


def test_rejected_options_raise_like_the_reference():
    """SURVEY 8b "Errors": unsupported user / doc models raise Exception('Unsupport ... model'), unknown scorers
    NotImplementedError, a 'dot' scorer over vectors of different widths ValueError (keras.layers.Dot.build) — what the
    reference's own build_model raised for each bad configuration (make_ref_golden.run_error_cases), and what the mirror
    raises before it ever touches the GPU"""
    d = tempfile.mkdtemp()
    synth.write_dataset(d, SH)
    cd = tempfile.mkdtemp()
    csh = mk.cook_shape()
    synth.write_cook_npz(cd, csh)
    ref = {str(a): (str(b), str(c)) for a, b, c in zip(GOLD['errors/labels'], GOLD['errors/kinds'], GOLD['errors/messages'])}
    for label, task_name, arch, score_model, extra in mk.ERROR_CASES:
        if task_name == 'Cook':
            cfg = settings.Config(dict(task='Cook', arch=arch, input_training_data_path=cd, days=30, window_size=csh.W, batch_size=8,
                                       title_filter_shape=(csh.F, 3), user_embedding_dim=csh.U, dropout=0.0, score_model=score_model,
                                       use_vertical=True, vertical_embedding_dim=mk.COOK_DV, subvertical_embedding_dim=mk.COOK_DS,
                                       precision='fp32', **extra))
        else:
            cfg = settings.Config(dict(task=task_name, arch=arch, score_model=score_model, input_training_data_path=d,
                                       title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B,
                                       textual_embedding_dim=SH.E, title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U,
                                       debug=True, dropout=0.0, precision='fp32', **extra))
        kind, msg = ref[label]
        assert kind != 'none', label
        with pytest.raises(Exception) as info:
            task.get(cfg).build_model(0)
        assert type(info.value).__name__ == kind, (label, type(info.value).__name__, kind, str(info.value))
        if kind == 'Exception':
            assert str(info.value) == msg, (label, str(info.value), msg)
