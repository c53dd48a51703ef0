"""GPU tests of the reference-facing surface (task/paper.py mirror + Keras Model protocol) and of the committed
golden fixtures, through the C-ABI."""
import os
import tempfile

import numpy as np
import pytest
import torch

from mnexp_b200 import settings, synth, task
from oracle import lstur_numpy as on

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lstur_golden.npz'))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('arch', ['igru', 'gru', 'hgru', 'nigru'])
def test_engine_matches_golden_fixtures(lib, arch):
    from mnexp_b200.engine import LsturEngine
    sh = synth.SHAPES['tiny']
    tok = GOLD['doc_tokens']
    P = synth.make_weights(sh, arch=arch, bias_noise=0.05, seed=4242)
    b = {k: GOLD['%s/batch/%s' % (arch, k)] for k in ('user', 'hist_doc', 'cand_doc')}
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch=arch, doc_tokens=tok, lr=1e-3, sparse_user_adam=False)
    db = eng.to_device_batch(b)
    eng.forward(db, training=True)
    eng.backward(db)
    assert rel(eng.view('probs').reshape(sh.B, -1).cpu().numpy(), GOLD['%s/probs' % arch]) < 2e-5
    assert rel(eng.view('user_vec').reshape(sh.B, -1).cpu().numpy(), GOLD['%s/user_vec' % arch]) < 2e-5
    assert rel(eng.score_sigmoid().cpu().numpy(), GOLD['%s/sigmoid' % arch]) < 2e-5
    assert abs(eng.loss() - float(GOLD['%s/loss' % arch])) < 2e-5
    g = eng.get_grads_dict()
    for k in g:
        key = '%s/grad/%s' % (arch, k)
        if key in GOLD:
            assert rel(g[k], GOLD[key]) < 5e-5, k
    # two dense Keras-Adam steps
    eng2 = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch=arch, doc_tokens=tok, lr=1e-3, sparse_user_adam=False)
    losses = [float(eng2.train_step(db)[0]) for _ in range(2)]
    assert np.abs(np.array(losses) - GOLD['%s/adam_losses' % arch]).max() < 1e-4
    assert np.abs(eng2.get_weights_dict()['conv_w'] - GOLD['%s/adam_conv_w' % arch]).max() < 2e-5


def _handler(arch='igru', task_name='Seq2VecPaperSoftmaxId', precision='fp32', batch_size=8, **kw):
    sh = synth.SHAPES['tiny']
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sh)
    cfg = settings.Config(dict(task=task_name, arch=arch, input_training_data_path=d, title_shape=sh.L,
                               window_size=sh.W, negative_samples=sh.K, batch_size=batch_size, textual_embedding_dim=sh.E,
                               title_filter_shape=(sh.F, 3), user_embedding_dim=sh.U, debug=True, dropout=0.0,
                               precision=precision, validation_impression=5, testing_impression=5, epochs=2, **kw))
    return sh, task.get(cfg)


def _oracle_params(model):
    names = [k for k in model.WEIGHT_ORDER if k in model._current()]
    P = dict(zip(names, model.get_weights()))
    P['att_w'] = P['att_w'].reshape(-1)
    return P


@pytest.mark.parametrize('task_name,arch', [('Seq2VecPaperSoftmaxId', 'igru'), ('Seq2VecPaperSoftmaxId', 'gru'),
                                            ('Seq2VecPaperSoftmax', 'gru')])
def test_model_builder_surface(lib, task_name, arch):
    sh, h = _handler(arch, task_name)
    model = h.build_model(0)
    assert model is h.model and h.test_model is not None and h.build_model(1) is model
    x, y = next(h.train)
    # predict == oracle on the same weights (the reference's forward semantics), float64 token arrays accepted
    P = _oracle_params(model)
    user = x[0] if h.HAS_USER else np.zeros(len(y), dtype=int)
    clicked = x[1] if h.HAS_USER else x[0]
    cands = np.stack(x[2:] if h.HAS_USER else x[1:], 1)
    ref = on.lstur_forward(P, user, clicked.astype(int), cands.astype(int), arch=h._engine_arch(), aux=True)
    assert rel(model.predict(x), ref['probs']) < 2e-5
    one = (x[:2] if h.HAS_USER else x[:1]) + [x[2 if h.HAS_USER else 1]]
    s = h.test_model.predict(one)
    assert s.shape == (len(y), 1) and rel(s[:, 0], ref['sigmoid'][:, 0]) < 2e-5
    ev = model.evaluate(x, y)
    assert abs(ev[0] - on.categorical_crossentropy(y, ref['probs'])) < 1e-5 and model.metrics_names == ['loss', 'categorical_accuracy']
    # fit_generator trains: the loss on a fixed batch goes down
    l0 = model.evaluate(x, y)[0]
    hist = model.fit_generator(h.train, 12, epochs=1, initial_epoch=0, verbose=0)
    assert set(hist.history) == {'loss', 'categorical_accuracy'} and len(hist.history['loss']) == 1
    for _ in range(15):
        model.train_on_batch(x, y)
    assert model.evaluate(x, y)[0] < l0
    # get_weights / set_weights round trip
    w = model.get_weights()
    model.set_weights([a * 0 + 0.01 for a in w])
    assert abs(model.get_weights()[1] - 0.01).max() < 1e-7
    model.set_weights(w)
    assert rel(model.get_weights()[1], w[1]) == 0
    # json + pkl model files (utils.py:66-79): the pkl is a plain pickled list of numpy arrays
    from mnexp_b200 import utils as mutils
    import pickle
    paths = (os.path.join(h.config.input_training_data_path, 'm.json'), os.path.join(h.config.input_training_data_path, 'm.pkl'))
    mutils.save_model(paths, model)
    with open(paths[1], 'rb') as f:
        raw = pickle.load(f)
    assert isinstance(raw, list) and all(isinstance(a, np.ndarray) for a in raw) and len(raw) == len(w)
    model.set_weights([a * 0 for a in w])
    loaded = mutils.load_model(paths)
    assert loaded.config['arch'] == h._engine_arch() and loaded.config['weight_names'][0] == 'word_emb'
    loaded.apply_to(model)
    assert all(np.array_equal(a, b) for a, b in zip(model.get_weights(), w))
    # callback: LR decay + ranking metrics over validation impressions (task/paper.py:497-524)
    lr0 = model.optimizer.lr.value
    h.callback(0)
    assert abs(h.model.optimizer.lr.value - lr0 * h.config.learning_rate_decay) < 1e-12
    assert 0.0 <= h.last_evaluation['auc'] <= 1.0 and h.model is model
    # doc_encoder layer by name (task/test_pipeline.py:28)
    dv = model.get_layer('doc_encoder').predict(clicked[0])
    assert rel(dv, on.news_encoder(clicked[0].astype(int), _oracle_params(model))) < 5e-5


def test_surface_tensor_core_precision(lib):
    sh = synth.Shape('t16', 50, 80, 120, L=7, W=5, K=2, B=6, E=12, F=16, U=8)
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sh)
    cfg = settings.Config(dict(task='Seq2VecPaperSoftmaxId', arch='igru', input_training_data_path=d, title_shape=sh.L,
                               window_size=sh.W, negative_samples=sh.K, batch_size=8, textual_embedding_dim=sh.E,
                               title_filter_shape=(sh.F, 3), user_embedding_dim=sh.U, debug=True, dropout=0.2))
    h = task.get(cfg)
    model = h.build_model(0)
    assert h._core.precision() == 'fp16_tc'
    x, y = next(h.train)
    P = _oracle_params(model)
    ref = on.lstur_forward(P, x[0], x[1].astype(int), np.stack(x[2:], 1).astype(int), arch='igru')
    assert rel(model.predict(x), ref) < 1e-3
    hist = model.fit_generator(h.train, 5, epochs=1)
    assert np.isfinite(hist.history['loss'][0])


@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
@pytest.mark.parametrize('arch', ['igru', 'gru'])
def test_decomposed_pipeline_matches_full_model(lib, arch, precision):
    from mnexp_b200.engine import LsturEngine
    """The reference's own self-check (task/test_pipeline.py:257-265): scoring through cached document vectors
    (doc_encoder once per document, then user_encoder + dot) equals the full model on the same impressions."""
    sh = synth.SHAPES['C1']
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch=arch, bias_noise=0.05)
    (b,), _ = synth.make_batches(sh, 1)
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch=arch, doc_tokens=tok, precision=precision, training=False)
    db = eng.to_device_batch(b)
    full = eng.forward(db).cpu().numpy().copy()
    full_sig = eng.score_sigmoid().cpu().numpy().copy()
    table = eng.build_doc_table()
    assert table.shape == (tok.shape[0], eng.D) and float(table[0].abs().max()) == 0.0
    dec = eng.forward_docvecs(db, table).cpu().numpy().copy()
    dec_sig = eng.score_sigmoid().cpu().numpy()
    # the same kernels encode a document wherever its title sits in a batch: fp32 results agree to rounding of the
    # differently-tiled GEMMs, the tensor-core path bit for bit in the encoder
    tol = 1e-5
    assert np.abs(dec - full).max() < tol and np.abs(dec_sig - full_sig).max() < tol
    # and against the float64 oracle through cached vectors
    ref = on.lstur_forward(P, b['user'], tok[b['hist_doc']], tok[b['cand_doc']], arch=arch)
    assert np.abs(dec - ref).max() / np.abs(ref).max() < (2e-5 if precision == 'fp32' else 1e-3)


def test_pipeline_files_mirror(lib):
    """TestPipeline over the reference's three pipeline files (docs.tsv / UserClick.tsv / userDocPair.tsv):
    test_correct's two printed numbers agree, and every written score is the dot product of the cached vectors."""
    from mnexp_b200.task.test_pipeline import TestPipeline
    sh, h = _handler('igru', pipeline_input=tempfile.mkdtemp(), name='t')
    h.build_model(0)
    cfg = h.config
    g = np.random.default_rng(3)
    docs = sorted(k for k in h.docs if k != 0)[:40]
    with open(cfg.pipeline_inputs[0], 'w') as f:
        for d in docs:
            toks = [int(x) for x in h.docs[d].title if x != 0] or [1]
            f.write('d%d\t%s\n' % (d, ' '.join(map(str, toks))))
    users = []
    with open(cfg.pipeline_inputs[1], 'w') as f:
        for u in range(11):
            clicks = ['d%d' % docs[i] for i in g.integers(0, len(docs), g.integers(1, sh.W + 3))]
            if u == 3:
                clicks.append('d_unknown')                  # not in docs.tsv: stays a zero vector
            users.append((str(u), 'x', clicks))
            f.write('%d\tx\t%s\n' % (u, '#N#'.join(clicks)))
    with open(cfg.pipeline_inputs[2], 'w') as f:
        for u in range(11):
            for d in g.integers(0, len(docs), 3):
                f.write('%d\tx\td%d\n' % (u, docs[d]))
        f.write('999\tx\td%d\n' % docs[0])                   # unknown user: skipped like the reference does
    tp = TestPipeline(cfg)
    tp.load_model(h)
    tp.test_doc_vec()
    tp.test_user_vec()
    tp.test_user_doc_score()
    assert len(tp.doc_vec) == len(docs) and len(tp.user_vec) == 11
    lines = [l.rstrip('\n').split('\t') for l in open(cfg.pipeline_output)]
    assert len(lines) == 33
    for uid, ut, d, sc in lines:
        assert abs(float(sc) - float(np.dot(tp.user_vec[uid + ut].astype(np.float64), tp.doc_vec[d].astype(np.float64)))) < 1e-6
    pred, sigm = tp.test_correct()
    assert abs(float(np.asarray(pred).reshape(-1)[0]) - float(np.asarray(sigm).reshape(-1)[0])) < 1e-5


def test_gpu_ranking_metrics_match_host(lib):
    """AUC / nDCG@10 / nDCG@5 / MRR per impression on the GPU == the reference's host formulas (utils.py:106-124,
    sklearn roc_auc_score) on ragged impressions, including tied scores (same tie order: descending index)."""
    from sklearn.metrics import roc_auc_score
    from mnexp_b200 import metrics, utils
    g = np.random.default_rng(9)
    scores, labels = [], []
    for i in range(200):
        c = int(g.integers(2, 300))
        s = g.random(c).astype(np.float32)
        if i % 7 == 0:
            s = np.round(s, 1)                                  # many ties
        y = (g.random(c) < 0.15).astype(np.float32)
        y[g.integers(0, c)] = 1.0
        y[(np.argmax(y) + 1) % c] = 0.0
        scores.append(s); labels.append(y)
    got = metrics.ranking_metrics(scores, labels)
    for i, (s, y) in enumerate(zip(scores, labels)):
        order = np.lexsort((-np.arange(len(s)), -s.astype(np.float64)))        # descending score, ties by descending index
        yo = y[order].astype(np.float64)
        disc = np.log2(np.arange(len(s)) + 2)
        ideal = np.sort(y.astype(np.float64))[::-1]
        ndcg = lambda k: np.sum((2 ** yo[:k] - 1) / disc[:k]) / np.sum((2 ** ideal[:k] - 1) / disc[:k])
        mrr = np.sum(yo / (np.arange(len(s)) + 1)) / np.sum(yo)
        want = np.array([roc_auc_score(y, s), ndcg(10), ndcg(5), mrr])
        assert np.abs(got[i] - want).max() < 2e-6, (i, got[i], want)
        if i % 7 != 0:                                          # without ties this is exactly the reference's utils code
            assert abs(utils.ndcg_score(y, s, 10) - got[i, 1]) < 2e-6 and abs(utils.mrr_score(y, s) - got[i, 3]) < 2e-6


def test_device_batch_assembly_matches_window(lib):
    """lstur_assemble_batch: history windows bit-exact with Seq2Vec.Window replayed on the host (task/seq2vec.py:17-53,
    task/paper.py:396-405), positive = the click, negatives drawn from the impression's negatives with replacement."""
    from mnexp_b200.assemble import DeviceBatcher
    sh, h = _handler('igru')
    B, W, K = 32, sh.W, sh.K
    bt = DeviceBatcher(h.data, B, W, K, seed=5)
    # host replay of train_gen: (user, window ids, pos, impression negatives) per sample, in generator order
    want = []
    for user, (ih, _) in enumerate(h.data):
        ch = h.Window(h.docs, W)
        for imp in ih:
            for pos in imp.pos:
                if ch.count:
                    want.append((user, ch.get_ids().copy(), pos, set(imp.neg)))
                ch.push(pos)
    assert bt.n_samples == len(want) and len(want) >= B
    draws = {}
    for rep in range(3):
        idx = torch.as_tensor(np.random.default_rng(rep).integers(0, len(want), B).astype(np.int32)).cuda()
        db = bt.assemble(idx)
        user, hist, cand = (db[k].cpu().numpy() for k in ('user', 'hist_doc', 'cand_doc'))
        for b, s in enumerate(idx.cpu().numpy()):
            u, win, pos, negs = want[s]
            assert user[b] == u and np.array_equal(hist[b], win) and cand[b, 0] == pos
            assert all(int(x) in negs for x in cand[b, 1:])
            draws.setdefault(int(s), []).extend(int(x) for x in cand[b, 1:])
    # with replacement: over all draws more than one distinct negative is used wherever the impression offers several
    multi = [s for s in draws if len(want[s][3]) > 3 and len(draws[s]) >= 8]
    assert not multi or any(len(set(draws[s])) > 1 for s in multi)
    # the assembled batch trains
    eng = h._core.engine_train(B) if hasattr(h, '_core') else None
    if eng is None:
        h.build_model(0)
        eng = h._core.engine_train(B)
    loss = float(eng.train_step(bt.next_batch())[0])
    assert np.isfinite(loss)


@pytest.mark.parametrize('arch,score_model', [('ngru', 'dnn'), ('dgru', 'ddot'), ('niavg', 'ddot'), ('igru', 'dnn'),
                                               ('iigru', 'dot')])
def test_model_builder_scorers_and_concat_archs(lib, arch, score_model):
    """--score-model dnn / ddot (task/paper.py:448-455) with the 2U-wide 'ngru' / 'dgru' user vectors and 'niavg',
    through the reference's task-handler surface."""
    sh, h = _handler(arch, 'Seq2VecPaperSoftmaxId', score_model=score_model)
    model = h.build_model(0)
    x, y = next(h.train)
    P = _oracle_params(model)
    for k in ('so_w',):
        if k in P:
            P[k] = P[k].reshape(-1, 1)
    cands = np.stack(x[2:], 1)
    ref = on.lstur_forward(P, x[0], x[1].astype(int), cands.astype(int), arch=arch, score_model=score_model, aux=True)
    assert rel(model.predict(x), ref['probs']) < 2e-5
    s = h.test_model.predict(x[:2] + [x[2]])
    assert s.shape == (len(y), 1) and rel(s[:, 0], ref['sigmoid'][:, 0]) < 2e-5
    l0 = model.evaluate(x, y)[0]
    for _ in range(25):
        model.train_on_batch(x, y)
    assert model.evaluate(x, y)[0] < l0


def test_concat_archs_need_a_dense_scorer(lib):
    with pytest.raises(ValueError):
        _handler('ngru', 'Seq2VecPaperSoftmaxId', score_model='dot')[1].build_model(0)


@pytest.mark.parametrize('task_name,arch,oarch,score', [
    ('Seq2VecPaperId', 'igru', 'igru', 'dnn'), ('Seq2VecPaperId', 'gru', 'ngru', 'dnn'),
    ('Seq2VecPaperId', 'iigru', 'iicat', 'dnn'), ('Seq2VecPaperDot', 'gru', 'nigru', 'dot'),
    ('Seq2VecPaper', 'avg', 'niavg', 'dnn'),
])
def test_sigmoid_family_surface(lib, task_name, arch, oarch, score):
    """Seq2VecPaper / Seq2VecPaperDot / Seq2VecPaperId (task/paper.py:6-383): (history, candidate, label) samples,
    sigmoid score, weighted BCE; predict == oracle, evaluate == Seq2Vec.loss, training reduces the loss, test yields
    per-impression (scores, labels)."""
    from oracle import lstur_torch as ot
    import torch
    sh, h = _handler(arch, task_name, gain=1.5)
    model = h.build_model(0)
    x, y = next(h.train)
    assert y.shape == (8,) and len(x) == (3 if h.HAS_USER else 2)
    P = _oracle_params(model)
    if 'so_w' in P:
        P['so_w'] = P['so_w'].reshape(-1, 1)
    user = x[0] if h.HAS_USER else np.zeros(len(y), dtype=int)
    clicked, cand = (x[1], x[2]) if h.HAS_USER else (x[0], x[1])
    ora = ot.LsturOracle(P, arch=oarch, score_model=score)
    u, c, d = ora._ints(user, clicked.astype(int), cand.astype(int)[:, None])
    ref = ot.forward(ora.P, u, c, d, arch=oarch, score_model=score, head='sigmoid').detach()
    p = model.predict(x)
    assert p.shape == (8, 1) and rel(p, ref.numpy()) < 2e-5
    ev = model.evaluate(x, y)
    l_ref = float(ot.weighted_bce(torch.tensor(y, dtype=torch.float64)[:, None], ref, gain=1.5, negative_samples=sh.K))
    assert abs(ev[0] - l_ref) < 1e-5 and model.metrics_names == ['loss', 'auc_roc']
    for _ in range(25):
        out = model.train_on_batch(x, y)
    assert len(out) == 2 and model.evaluate(x, y)[0] < ev[0]
    scores, labels = next(h.test)
    assert scores.shape == labels.shape and scores.ndim == 1 and set(np.unique(labels)) <= {0, 1}


def _cook_npz(d, sh, n_train=24, n_test=10, seed=0):
    """train/test .npz in the reference's cook layout (task/cook.py:14-28) + Vocab.tsv.npy."""
    synth.write_cook_npz(d, sh, n_train, n_test, seed)


@pytest.mark.parametrize('arch,oarch,score_model,vtype', [('ingru', 'igru', 'ddot', 'vs'), ('igru', 'ngru', 'dnn', 'v'),
                                                          ('inigru', 'iicat', 'ddot', 's'), ('avg', 'niavg', 'dnn', 'vs'),
                                                          # the remaining Cook.get_user_encoder branches (task/cook.py:155-193)
                                                          ('iavg', 'iavg', 'dnn', 'vs'), ('iatt', 'iatt', 'ddot', 'v'),
                                                          ('ilstm', 'ilstm', 'dnn', 's'), ('inagru', 'inagru', 'ddot', 'vs'),
                                                          ('atgru', 'atgru', 'dnn', 'vs'), ('algru', 'algru', 'ddot', 'v')])
def test_cook_handler(lib, arch, oarch, score_model, vtype):
    """Cook (task/cook.py): .npz protocol with per-slot vertical / subvertical ids, [title ‖ Vemb ‖ Semb] news vectors,
    idx_mask on the user id embedding, linear 'ddot'; train_model / test_model through main.py's cook calls."""
    from oracle import lstur_torch as ot
    import torch
    sh = synth.SHAPES['tiny']
    d = tempfile.mkdtemp()
    _cook_npz(d, sh)
    cfg = settings.Config(dict(task='Cook', arch=arch, input_training_data_path=d, days=30, window_size=sh.W,
                               batch_size=8, title_filter_shape=(sh.F, 3), user_embedding_dim=sh.U, dropout=0.0,
                               score_model=score_model, use_vertical=True, use_vertical_type=vtype,
                               vertical_embedding_dim=3, subvertical_embedding_dim=5, precision='fp32', validation_step=6,
                               lrd_on_epochs=[0]))
    h = task.get(cfg)
    model = h.build_model(0)
    x, y = h.train()
    P = {k: np.asarray(v) for k, v in model.get_weights_dict().items()}
    ora = ot.LsturOracle(P, arch=oarch, score_model=score_model)
    n = len(y[0])
    kw = dict(arch=oarch, score_model=score_model, u0_scale=torch.tensor(x[1], dtype=torch.float64).reshape(n, 1))
    if 'vert_emb' in P:
        kw.update(hist_vert=x[3], cand_vert=x[6])
    if 'subvert_emb' in P:
        kw.update(hist_subvert=x[4], cand_subvert=x[7])
    u, c, dd = ora._ints(x[0].reshape(-1), x[2], x[5])
    s = ot.score(*[ot.forward(ora.P, u, c, dd, aux=True, **kw)[k] for k in ('user_vec', 'cand_vec')], ora.P, score_model,
                 flavour='cook')
    ref = torch.softmax(s, -1).detach().numpy()
    assert rel(model.predict(x), ref) < 5e-5
    l0 = model.evaluate(x, y)[0]
    hist = model.fit(x, y, 8, epochs=3, initial_epoch=0, shuffle=True)
    assert len(hist.history['loss']) == 3 and model.evaluate(x, y)[0] < l0
    lr0 = model.optimizer.lr.value
    h.callback(0)
    assert abs(model.optimizer.lr.value - lr0 * cfg.learning_rate_decay) < 1e-12
    ev = h.test_model.evaluate(*h.valid())
    assert len(ev) == 2 and h.test_model.metrics_names == ['loss', 'auc_roc'] and np.isfinite(ev[0])
    feats, (users, imprs, mask, y_true) = h.test()
    pred = h.test_model.predict(feats).reshape(-1)
    assert pred.shape == y_true.shape and np.all((pred > 0) & (pred < 1))


def test_days_id_vert_surface(lib):
    """Seq2VecPaperSoftmaxDaysIdVert (task/paper.py:1138-1255): vertical ids ride next to the titles; news vector =
    [Dense(U)(title) ‖ Vemb[vertical]]; user encoder widened by vertical_embedding_dim."""
    sh, h = _handler('igru', 'Seq2VecPaperSoftmaxDaysIdVert', days=100000, vertical_embedding_dim=4)
    model = h.build_model(0)
    x, y = next(h.train)
    C = 1 + sh.K
    assert len(x) == 3 + 2 * C and x[2].shape == (8, sh.W) and x[3 + C].shape == (8,)
    assert set(np.unique(x[2])) <= set(range(16)) and (x[2][x[1].any(-1) == 0] == 0).all()     # pad slots: vertical 'N/A'
    P = _oracle_params(model)
    P['vert_emb'] = model._current()['vert_emb']
    assert P['vert_emb'].shape == (16, 4) and P['gru_wh'].shape[0] == sh.U + 4
    cands, cverts = np.stack(x[3:3 + C], 1), np.stack(x[3 + C:], 1)
    ref = on.lstur_forward(P, x[0], x[1].astype(int), cands.astype(int), arch='igru', aux=True, hist_vert=x[2],
                           cand_vert=cverts)
    assert rel(model.predict(x), ref['probs']) < 2e-5
    s = h.test_model.predict(x[:3] + [x[3], x[3 + C]])
    assert s.shape == (8, 1) and rel(s[:, 0], ref['sigmoid'][:, 0]) < 2e-5
    l0 = model.evaluate(x, y)[0]
    for _ in range(25):
        model.train_on_batch(x, y)
    assert model.evaluate(x, y)[0] < l0
    h.callback(0)            # swaps in test_model and ranks the validation impressions (task/paper.py:497-524)
    assert 0.0 <= h.last_evaluation['auc'] <= 1.0 and h.model is model
    dv = model.get_layer('doc_encoder').predict(x[1][0])
    assert dv.shape == (sh.W, sh.U)


def test_vertsup_surface(lib):
    """Seq2VecPaperSoftmaxDaysIdVertSup (task/paper.py:884-1000): two targets per batch, two-output model with
    loss = CE_ranking + gain * CE_vert, Keras' multi-output metric names; test_model is the plain scorer."""
    from oracle import lstur_torch as ot
    import torch
    sh, h = _handler('igru', 'Seq2VecPaperSoftmaxDaysIdVertSup', days=100000, hidden_dim=12, gain=0.5)
    model = h.build_model(0)
    x, y = next(h.train)
    C = 1 + sh.K
    assert len(x) == 2 + C and len(y) == 2 and y[0].shape == (8, C) and y[1].shape == (8, sh.W + C, 16)
    assert np.allclose(y[1].sum(-1), 1.0) and (y[1][:, :sh.W].argmax(-1)[x[1].any(-1) == 0] == 0).all()   # pad slots: 'N/A'
    assert model.metrics_names == ['loss', 'ranking_loss', 'vert_loss', 'ranking_categorical_accuracy', 'vert_categorical_accuracy']
    P = {k: np.asarray(v) for k, v in h._core.engine_train(8).get_weights_dict().items()}
    assert P['vs_w1'].shape == (sh.U, 12) and P['vs_w2'].shape == (12, 16)
    ora = ot.LsturOracle(P, arch='igru')
    u, c, d = ora._ints(x[0], x[1], np.stack(x[2:], 1))
    ids = y[1].argmax(-1)
    main, aux, vp = ot.loss_fn(ora.P, u, c, d, label=torch.tensor(y[0], dtype=torch.float64), arch='igru',
                               vert_labels=(ids[:, :sh.W], ids[:, sh.W:]), parts=True)
    ev = model.evaluate(x, y)
    assert abs(ev[1] - float(main)) < 1e-5 and abs(ev[2] - float(aux)) < 1e-5 and abs(ev[0] - float(main + 0.5 * aux)) < 1e-5
    pr = model.predict(x)
    assert pr[0].shape == (8, C) and pr[1].shape == (8, sh.W + C, 16) and rel(pr[1], vp.detach().numpy()) < 2e-5
    for _ in range(30):
        out = model.train_on_batch(x, y)
    ev2 = model.evaluate(x, y)
    assert len(out) == 5 and ev2[0] < ev[0] and ev2[2] < ev[2]
    s = h.test_model.predict(x[:2] + [x[2]])
    assert s.shape == (8, 1) and np.all((s > 0) & (s < 1))
    h.callback(0)
    assert 0.0 <= h.last_evaluation['auc'] <= 1.0


def test_vertalt_surface(lib):
    """Seq2VecPaperSoftmaxDaysIdVertAlt (task/paper.py:1003-1136): epochs x round; vertical model on 10 % of the documents;
    callback_valid alternates self.model between vert_model and seq_model; both share the doc encoder."""
    sh, h = _handler('igru', 'Seq2VecPaperSoftmaxDaysIdVertAlt', days=100000, round=3, batch_size=4)
    assert h.config.epochs == 2 * 3
    model = h.build_model(0)
    assert model is h.vert_model and h.train_seq is False
    nd = len(h.data_titles)
    assert len(h.train_index) == nd // 10 and len(h.valid_index) == nd - nd // 10
    assert h.training_step == len(h.train_index) // 4 and h.validation_step == len(h.valid_index) // 4
    train = h.train
    x, y = next(train)
    assert x.shape == (4, sh.L) and y.shape == (4, len(h.verticals)) and model.metrics_names == ['loss', 'categorical_accuracy']
    conv0 = h._core.engine_train(4).get_weights_dict()['conv_w'].copy()
    l0 = model.evaluate(x, y)[0]
    for _ in range(30):
        model.train_on_batch(x, y)
    assert model.evaluate(x, y)[0] < l0
    conv1 = h._core.train_engine.get_weights_dict()['conv_w']
    assert np.abs(conv1 - conv0).max() > 0                       # the vertical model trains the shared doc encoder
    ev = model.evaluate_generator(h.valid, 2)
    assert len(ev) == 2 and np.isfinite(ev[0])
    h.callback(0)
    h.callback_valid(0)                                          # epoch 0 of round 3: stay on the vertical model
    assert h.train_seq is False and h.build_model(1) is h.vert_model
    lr0 = model.optimizer.lr.value
    h.callback(1)                                                # epoch % round == round - 2: decay the vertical lr
    assert abs(model.optimizer.lr.value - lr0 * h.config.learning_rate_decay) < 1e-12
    h.callback_valid(1)
    assert h.train_seq is True and h.build_model(2) is h.seq_model and h.training_step == h.config.training_step
    xs, ys = next(train)                                         # the same generator now yields click batches
    assert len(xs) == 2 + 1 + sh.K and ys.shape == (4, 1 + sh.K)
    before = h.seq_model.evaluate(xs, ys)[0]
    for _ in range(25):
        h.seq_model.train_on_batch(xs, ys)
    assert h.seq_model.evaluate(xs, ys)[0] < before
    h.callback_valid(2)
    assert h.train_seq is False and h.model is h.vert_model


@pytest.mark.parametrize('task_name,arch,oarch', [('Seq2VecPaperSoftmax', 'att', 'att'), ('Seq2VecPaperSoftmax', 'avg', 'niavg'),
                                                  ('Seq2VecPaper', 'att', 'att')])
def test_attention_and_average_user_encoders_of_the_non_id_classes(lib, task_name, arch, oarch):
    """Seq2VecPaper.get_user_encoder (task/paper.py:199-221, inherited by Seq2VecPaperSoftmax): 'att' / 'avg' / 'gru'."""
    sh, h = _handler(arch, task_name)
    model = h.build_model(0)
    x, y = next(h.train)
    P = _oracle_params(model)
    if 'uatt_w' in P:
        P['uatt_w'] = P['uatt_w'].reshape(-1)
    if task_name == 'Seq2VecPaperSoftmax':
        ref = on.lstur_forward(P, np.zeros(len(y), dtype=int), x[0].astype(int), np.stack(x[1:], 1).astype(int), arch=oarch)
        assert rel(model.predict(x), ref) < 2e-5
    else:
        ref = on.lstur_forward(P, np.zeros(len(y), dtype=int), x[0].astype(int), x[1][:, None].astype(int), arch=oarch,
                               score_model='dnn', aux=True)
        assert rel(model.predict(x).reshape(-1), ref['sigmoid'].reshape(-1)) < 2e-5
    l0 = model.evaluate(x, y)[0]
    for _ in range(25):
        model.train_on_batch(x, y)
    assert model.evaluate(x, y)[0] < l0


def test_engine_matches_variant_golden_fixtures(lib):
    """Committed fixtures of the remaining archs / scorers / sigmoid family (tests/golden/make_golden_variants.py)."""
    import importlib.util
    from mnexp_b200.engine import LsturEngine
    here = os.path.dirname(__file__)
    spec = importlib.util.spec_from_file_location('mgv', os.path.join(here, 'golden', 'make_golden_variants.py'))
    mgv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgv)
    gold = np.load(os.path.join(here, 'golden', 'lstur_golden_variants.npz'))
    sh = synth.SHAPES['tiny']
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    eng_arch = {'iicat': 'iigru'}
    for arch, score_model, head in mgv.CASES:
        key = '%s-%s-%s' % (arch, score_model, head)
        P = synth.make_weights(sh, arch=arch, bias_noise=0.05, seed=5150, score_model=score_model)
        (b,), _ = synth.make_batches(sh, 1, seed=51)
        b = dict(b)
        kw = {}
        if head == 'bce':
            b['cand_doc'] = b['cand_doc'][:, :1]
            b['label'] = gold[key + '/label']
            kw = dict(flavour='sigmoid', loss='bce', gain=mgv.GAIN, bce_neg=sh.K)
        eng = LsturEngine(P, sh.B, sh.W, b['cand_doc'].shape[1], sh.L, arch=eng_arch.get(arch, arch), doc_tokens=tok,
                          score_model=score_model, **kw)
        db = eng.to_device_batch(b)
        probs = eng.forward(db, training=True, seed=1).cpu().numpy().copy()
        eng.backward(db)
        torch.cuda.synchronize()
        assert rel(probs, gold[key + '/probs']) < 5e-5, key
        assert abs(eng.loss() - float(gold[key + '/loss'])) < 5e-5 * max(1.0, abs(float(gold[key + '/loss']))), key
        g = eng.get_grads_dict()
        for k in g:
            gk = key + '/grad/' + k
            if gk in gold.files and k not in ('att_b', 'so_b'):
                assert rel(g[k], gold[gk]) < 1e-4, (key, k)


def test_named_encoder_layers_and_per_impression_scoring(lib):
    """get_layer('doc_encoder') / get_layer('user_encoder') with the input 'user_clicked_vec' (task/paper.py:160, 590, 632;
    task/test_pipeline.py:27-35, 87-88) compose to the full model; Seq2Vec.test's one-impression predict (same history on
    every row, task/seq2vec.py:202-206) encodes the history once and returns what row-by-row prediction returns."""
    sh, h = _handler('igru', 'Seq2VecPaperSoftmaxId')
    model = h.build_model(0)
    x, y = next(h.train)
    user, clicked, cands = x[0], x[1], np.stack(x[2:], 1)
    P = _oracle_params(model)
    ref = on.lstur_forward(P, user, clicked.astype(int), cands.astype(int), arch='igru', aux=True)
    ue, de = model.get_layer('user_encoder'), h.test_model.get_layer('doc_encoder')
    assert ue is h.user_encoder and ue.name == 'user_encoder' and de.name == 'doc_encoder'
    assert ue.get_layer('user_clicked_vec').input_shape == (None, sh.W, sh.U)
    assert de.layers[0].input_shape[-1] == sh.L
    # doc vectors of the history through the doc_encoder layer, masked like ComputeMasking does (task/paper.py:644-645)
    B, W, L = clicked.shape
    dv = de.predict(clicked.reshape(B * W, L)).reshape(B, W, -1)
    dv = dv * (clicked != 0).any(-1)[..., None]
    assert rel(dv, ref['hist_vec']) < 5e-5
    uv = ue.predict([user, dv])
    assert uv.shape == (B, sh.U) and rel(uv, ref['user_vec']) < 5e-5
    # one impression: 7 candidates against the history of row 0
    n = 7
    cand = cands.reshape(-1, L)[:n]
    assert cand.shape[0] == n
    imp = [np.repeat(user[:1], n), np.repeat(clicked[:1], n, 0), cand]
    fast = h.test_model.predict(imp)
    slow = np.concatenate([h.test_model.predict([a[i:i + 1] for a in imp]) for i in range(n)])
    assert fast.shape == (n, 1) and rel(fast, slow) < 1e-6
    # more candidates than one row holds (32): the history is re-encoded once per group of 32
    n2 = 70
    cand2 = np.concatenate([cands.reshape(-1, L)] * 3)[:n2]
    imp2 = [np.repeat(user[:1], n2), np.repeat(clicked[:1], n2, 0), cand2]
    fast2 = h.test_model.predict(imp2)
    slow2 = np.concatenate([h.test_model.predict([a[i:i + 1] for a in imp2]) for i in range(0, n2, 9)])
    assert rel(fast2[::9], slow2) < 1e-6


def test_batch_size_change_keeps_optimizer_state(lib):
    """predict / callbacks between training steps must not reset Adam: the inference engines share the live
    training engine's weights, and a training engine rebuilt for another batch size adopts moments, step count and the
    dropout seed counter."""
    sh, h = _handler('igru', 'Seq2VecPaperSoftmaxId')
    model = h.build_model(0)
    x, y = next(h.train)
    half = [a[:4] for a in x], y[:4]
    model.train_on_batch(*half)                        # batch 4 != config.batch_size 8
    e4 = h._core.train_engine
    assert e4.B == 4 and e4.t == 1
    model.predict(x)                                   # must not rebuild the training engine
    h.test_model.predict([x[0], x[1], x[2]])
    assert h._core.train_engine is e4 and e4.t == 1
    v_before = e4.adam_v.clone()
    model.train_on_batch(x, y)                         # batch 8: rebuilt, state adopted
    e8 = h._core.train_engine
    assert e8 is not e4 and e8.B == 8 and e8.t == 2 and e8.step_seed == 2
    # moments carried, not zeroed: v_new = 0.999 v_old + 0.001 g^2 >= 0.999 v_old element-wise
    assert float(v_before.max()) > 0 and bool((e8.adam_v >= 0.998 * v_before).all())


def test_test_set_aggregation_on_device_metrics(lib):
    """main.py:224-297 aggregation with the per-impression metrics from the device kernel == with sklearn / numpy"""
    from mnexp_b200 import evaluation
    from test_host_logic import _host_metrics, _synthetic_scored_test_set
    users, imprs, mask, yt, yp = _synthetic_scored_test_set(3)
    a = evaluation.aggregate(users, imprs, mask, yt, yp)
    b = evaluation.aggregate(users, imprs, mask, yt, yp, metric_fn=_host_metrics)
    for key in ('user', 'impr', 'iv_user', 'oov_user'):
        for f in ('auc', 'mrr', 'ndcgv', 'ndcgx', 'pos', 'size'):
            assert abs(getattr(a[key], f) - getattr(b[key], f)) < 1e-5, (key, f)


def test_enable_pretrain_encoder_round_trip_and_freeze(lib):
    """--enable-pretrain-encoder (task/paper.py:103-107): encoder{name}.json/.pkl written from one model's doc_encoder is
    loaded into a new model; without --pretrain-encoder-trainable its weights stay put while the rest trains."""
    from mnexp_b200 import utils as mutils
    sh, h = _handler('igru', 'Seq2VecPaperSoftmaxId')
    m1 = h.build_model(0)
    x, y = next(h.train)
    for _ in range(3):
        m1.train_on_batch(x, y)
    enc = m1.get_layer('doc_encoder')
    d = h.config.input_training_data_path
    h.config.output_model_path = d
    mutils.save_model(h.config.encoder_output, enc)
    sh2, h2 = _handler('igru', 'Seq2VecPaperSoftmaxId', enable_pretrain_encoder=True, input_previous_model_path=d)
    m2 = h2.build_model(0)
    w1, w2 = enc.get_weights(), m2.get_layer('doc_encoder').get_weights()
    assert all(np.array_equal(a, b) for a, b in zip(w1, w2))
    dv1 = enc.predict(x[1][0])
    assert rel(m2.get_layer('doc_encoder').predict(x[1][0]), dv1) < 1e-6
    gru_before = h2._core.params['gru_wx'].copy()
    for _ in range(3):
        m2.train_on_batch(x, y)
    w3 = m2.get_layer('doc_encoder').get_weights()
    assert all(np.array_equal(a, b) for a, b in zip(w2, w3))                       # frozen encoder
    assert np.abs(h2._core.train_engine.get_weights_dict()['gru_wx'] - gru_before).max() > 1e-4   # the rest trains
    # --pretrain-encoder-trainable: the encoder moves too
    sh3, h3 = _handler('igru', 'Seq2VecPaperSoftmaxId', enable_pretrain_encoder=True, pretrain_encoder_trainable=True,
                       input_previous_model_path=d)
    m3 = h3.build_model(0)
    for _ in range(3):
        m3.train_on_batch(x, y)
    assert np.abs(m3.get_layer('doc_encoder').get_weights()[1] - w1[1]).max() > 1e-4
