"""Host-side logic (no GPU): reference data formats, Window / Impression semantics, generators, RNG replica."""
import os
import tempfile

import numpy as np
import pytest

from mnexp_b200 import document, rng, settings, synth, task, utils


def test_document_parsers():
    p = document.DocumentParser(document.parse_document(), document.pad_document(1, 5))
    assert p('3 7 9')[0].tolist() == [3, 7, 9, 0, 0]
    assert p('1 2 3 4 5 6 7')[0].tolist() == [1, 2, 3, 4, 5]          # truncation
    assert p('1 2#N#8 9')[0].tolist() == [1, 2, 0, 0, 0]             # size=1: first sentence only
    assert p('3 7 9').dtype == np.float64                            # reference feeds float64 token ids
    assert document.pad_document(2, 3)([[1], [], [2, 3]]).tolist() == [[1, 0, 0], [2, 3, 0]]


def _dataset(shape=None):
    sh = shape or synth.SHAPES['tiny']
    d = tempfile.mkdtemp()
    emb, tok = synth.write_dataset(d, sh)
    cfg = settings.Config(dict(task='Seq2VecPaperSoftmaxId', arch='igru', input_training_data_path=d,
                               title_shape=sh.L, window_size=sh.W, negative_samples=sh.K, batch_size=4,
                               textual_embedding_dim=sh.E, title_filter_shape=(sh.F, 3), user_embedding_dim=sh.U,
                               debug=True))
    return sh, d, emb, tok, cfg


def test_config_defaults_and_paths():
    c = settings.Config({})
    assert (c.learning_rate, c.learning_rate_decay, c.dropout, c.negative_samples) == (0.001, 0.2, 0.2, 4)
    assert c.title_filter_shape == (400, 3) and c.user_embedding_dim == 200 and c.score_model == 'dot'
    c2 = settings.Config(dict(node_count=3, input_training_data_path='/x', days=7, window_size=20))
    assert not hasattr(c2, 'node_count')                              # settings.py:97-99
    assert c2.training_data_input == '/x/ClickData.tsv' and c2.train_npz_input == '/x/train_7days_20window.npz'


def test_docmeta_vocab_roundtrip():
    sh, d, emb, tok, cfg = _dataset()
    h = task.get(cfg)
    assert np.array_equal(h.doc_token_table(), tok)
    assert np.all(h.docs[0].title == 0)
    e2 = utils.load_textual_embedding(os.path.join(d, 'Vocab.tsv'), sh.E)
    assert e2.shape == emb.shape and np.all(e2[0] == 0) and np.array_equal(e2[1:], emb[1:])


def test_window_and_generators():
    sh, d, emb, tok, cfg = _dataset()
    h = task.get(cfg)
    w = h.Window(h.docs, 3)
    assert w.get_ids().tolist() == [0, 0, 0] and w.count == 0
    for x in (5, 6, 7, 8):
        w.push(x)
    assert w.get_ids().tolist() == [6, 7, 8] and w.count == 4      # FIFO, left-padded
    assert np.array_equal(w.get_title(), tok[[6, 7, 8]].astype(np.float64))
    imp = h.Impression('3 4#TAB#9 10 11#TAB#01/05/2019 03:07:00 PM')
    assert imp.pos == [3, 4] and imp.neg == [9, 10, 11] and imp.time.hour == 15
    assert set(imp.negative_samples(50)) <= {9, 10, 11}
    g = h.train_gen()
    for _ in range(20):
        s = next(g)
        user, clicked, pos = s[0], s[1], s[2]
        assert len(s) == 2 + 1 + sh.K + 1 and s[-1] == [1] + [0] * sh.K
        assert clicked.shape == (sh.W, sh.L) and pos.shape == (sh.L,)
        assert (clicked != 0).any()                               # train samples need >= 1 click (ch.count)
        pad = ~(clicked != 0).any(-1)
        assert not pad[np.argmax(~pad):].any()                      # padding is on the left only
        assert 0 <= user < len(h.data)
    x, y = next(h.train)
    assert [a.shape for a in x] == [(4,), (4, sh.W, sh.L)] + [(4, sh.L)] * (1 + sh.K) and y.shape == (4, 1 + sh.K)
    for b in h.test_gen():
        assert all(len(r) == 4 for r in b) and sum(r[-1] for r in b) >= 1
        break


def test_unsupported_options_raise_like_the_reference():
    sh, d, emb, tok, cfg = _dataset()
    for kw, exc in ((dict(arch='nope'), Exception), (dict(news_encoder='gruatt'), Exception),
                    (dict(score_model='zzz'), NotImplementedError)):
        c = settings.Config(dict({k: getattr(cfg, k) for k in ('task', 'input_training_data_path', 'title_shape',
                                                                'window_size', 'negative_samples', 'batch_size',
                                                                'textual_embedding_dim', 'title_filter_shape',
                                                                'user_embedding_dim', 'debug', 'arch')}, **kw))
        h = task.get(c)
        try:
            h.build_model(0)
        except exc:
            pass
        else:
            raise AssertionError(kw)


def test_ranking_metrics_known_answers():
    y, s = np.array([0, 1, 0, 1]), np.array([0.1, 0.9, 0.8, 0.3])
    assert abs(utils.mrr_score(y, s) - (1 / 1 + 1 / 3) / 2) < 1e-12
    assert abs(utils.ndcg_score(y, s, 10) - (1 + 1 / np.log2(4)) / (1 + 1 / np.log2(3))) < 1e-12
    assert utils.ndcg_score(y, y.astype(float), 5) == 1.0


def test_rng_replica_is_deterministic_and_unbiased():
    a = rng.dropout_multiplier(3, 200000, 0.2)
    assert np.array_equal(a, rng.dropout_multiplier(3, 200000, 0.2))
    assert not np.array_equal(a, rng.dropout_multiplier(4, 200000, 0.2))
    assert abs((a == 0).mean() - 0.2) < 0.005 and set(np.unique(a)) == {0.0, float(np.float32(1) / (np.float32(1) - np.float32(0.2)))}
    assert rng.rng_u32(1, np.array([2 ** 40 + 5], dtype=np.uint64))[0] != rng.rng_u32(1, np.array([5], dtype=np.uint64))[0]


def test_synth_shapes_match_baseline_configs():
    c3 = synth.SHAPES['C3']
    assert (c3.n_users, c3.n_news, c3.vocab, c3.W, c3.B, c3.L, c3.K, c3.E, c3.F, c3.U) == \
           (1_000_000, 130_000, 100_000, 50, 1024, 30, 4, 300, 400, 200)
    c5 = synth.SHAPES['C5']
    assert (c5.W, c5.L, c5.B) == (200, 50, 2048)
    b, pad = synth.make_batches(synth.SHAPES['C1'], 2)
    assert b[0]['hist_doc'].shape == (64, 50) and 0.2 < pad < 0.8
    h = b[0]['hist_doc']
    assert np.all((h[:, 1:] != 0) | (h[:, :-1] == 0))             # left padding only


def test_days_window_and_sigmoid_generators():
    """Time-window variants (task/paper.py:668-792) and the sigmoid family's (history, candidate, label) samples."""
    from datetime import datetime
    sh, d, emb, tok, cfg = _dataset()
    mk = lambda **kw: task.get(settings.Config(dict({k: getattr(cfg, k) for k in (
        'input_training_data_path', 'title_shape', 'window_size', 'negative_samples', 'batch_size',
        'textual_embedding_dim', 'title_filter_shape', 'user_embedding_dim', 'debug')}, **kw)))
    h = mk(task='Seq2VecPaperSoftmaxDaysId', arch='igru', days=2)
    w = h.Window(h.docs, 3, 2)
    t = lambda day: datetime(2019, 1, day)
    assert w.count(t(1)) == 0 and w.get_ids(t(1)) == [0, 0, 0]
    w.push(5, t(1)); w.push(6, t(2)); w.push(7, t(4))
    assert w.get_ids(t(4)) == [0, 6, 7] and w.count(t(4)) == 2            # the day-1 click expired after 2 days
    assert w.get_ids(t(7)) == [0, 0, 0] and w.count(t(7)) == 0
    assert np.array_equal(w.get_title(t(4)), tok[[0, 6, 7]].astype(np.float64))
    # with an unbounded horizon the time-window generator equals the plain one (same negative-sampling stream)
    far = mk(task='Seq2VecPaperSoftmaxDaysId', arch='igru', days=100000)
    plain = mk(task='Seq2VecPaperSoftmaxId', arch='igru')
    np.random.seed(3); a = [next(g) for g in [far.train_gen()] for _ in range(30)]
    np.random.seed(3); b = [next(g) for g in [plain.train_gen()] for _ in range(30)]
    assert all(x[0] == y[0] and all(np.array_equal(p, q) for p, q in zip(x[1:-1], y[1:-1])) for x, y in zip(a, b))
    # a short horizon only removes history: every sample's live clicks are a subset of the plain window's
    short = mk(task='Seq2VecPaperSoftmaxDays', arch='gru', days=1)
    s = next(short.train_gen())
    assert len(s) == 1 + 1 + sh.K + 1 and (s[0] != 0).any()
    # sigmoid family: K negatives follow each positive, labels 1, 0, 0, ...
    sg = mk(task='Seq2VecPaperId', arch='igru')
    g = sg.train_gen()
    rows = [next(g) for _ in range(2 * (1 + sh.K))]
    assert [r[-1] for r in rows] == ([1] + [0] * sh.K) * 2 and all(len(r) == 4 for r in rows)
    assert all(np.array_equal(rows[0][1], r[1]) for r in rows[:1 + sh.K])     # one history per positive and its negatives
    x, y = next(sg.train)
    assert [a.shape for a in x] == [(4,), (4, sh.W, sh.L), (4, sh.L)] and y.shape == (4,)
    nd = mk(task='Seq2VecPaperDot', arch='gru')
    assert len(next(nd.train_gen())) == 3


def _synthetic_scored_test_set(seed=0):
    g = np.random.default_rng(seed)
    users, imprs, mask, yt, yp = [], [], [], [], []
    for u in range(12):
        mk = int(g.integers(0, 2))
        for im in range(int(g.integers(1, 4))):
            n = int(g.integers(3, 9))
            lab = np.zeros(n); lab[g.integers(0, n)] = 1
            users += [u] * n; imprs += [im] * n; mask += [mk] * n; yt += list(lab); yp += list(g.random(n))
    return tuple(map(np.asarray, (users, imprs, mask, yt, yp)))


def _host_metrics(S, Y):
    from sklearn.metrics import roc_auc_score
    from mnexp_b200 import utils
    return np.array([[roc_auc_score(y, s), utils.ndcg_score(y, s, 10), utils.ndcg_score(y, s, 5), utils.mrr_score(y, s)]
                     for s, y in zip(S, Y)])


def test_per_user_in_vocab_oov_aggregation_mirrors_main_loop():
    """mnexp_b200.evaluation.aggregate == a replay of the state machine of main.py:250-287 (an impression closes on the
    first row of the next one, the last impression of the file is never closed; users are filed by mask at that row)."""
    from sklearn.metrics import roc_auc_score
    from mnexp_b200 import evaluation, utils
    users, imprs, mask, yt, yp = _synthetic_scored_test_set()
    res = evaluation.aggregate(users, imprs, mask, yt, yp, metric_fn=_host_metrics)
    pu, pi, index, ii = users[0], imprs[0], 0, 0
    IR, UR, IV, OOV = [], [], [], []
    for i in range(1, len(yp)):
        u, im = users[i], imprs[i]
        if u != pu or im != pi:
            y, s = yt[index:i], yp[index:i]
            IR.append(dict(auc=roc_auc_score(y, s), mrr=utils.mrr_score(y, s), ndcgv=utils.ndcg_score(y, s, 5),
                           ndcgx=utils.ndcg_score(y, s, 10), pos=y.sum(), size=i - index, idx=len(IR)))
            index, pi = i, im
        if u != pu:
            avg = {k: np.mean([r[k] for r in IR[ii:]]) for k in IR[0]}
            UR.append(avg)
            (IV if mask[index] == 1 else OOV).append(avg)
            ii, pu = len(IR), u
    for key, lst in (('impr', IR), ('user', UR), ('iv_user', IV), ('oov_user', OOV)):
        for f in ('auc', 'mrr', 'ndcgv', 'ndcgx', 'pos', 'size', 'idx'):
            assert abs(getattr(res[key], f) - np.mean([r[f] for r in lst])) < 1e-9, (key, f)
    assert res['user'].info['num'] == res['user'].idx * 2 + 1 and set(res['impr'].result) == {'auc', 'ndcgx', 'ndcgv', 'mrr'}


def test_pretrained_encoder_files_keras_and_own_format(tmp_path):
    """utils.load_model reads (a) the json + pkl pair the REFERENCE writes for a doc encoder — json.dump of Keras'
    to_json() string, pickle of get_weights() (utils.py:66-79) — mapping the arrays by the layer classes of the json, and
    (b) the pair written by save_model here; unexpected graphs raise instead of guessing."""
    import json
    import pickle
    from mnexp_b200 import utils
    g = np.random.default_rng(0)
    V, E, F, U = 20, 12, 16, 8
    W = [g.standard_normal((V, E)), g.standard_normal((3, E, F)), g.standard_normal(F), g.standard_normal((F, 1)), g.standard_normal(1),
         g.standard_normal((F, U)), g.standard_normal(U)]
    W = [a.astype(np.float32) for a in W]
    layer = lambda cls, name: dict(class_name=cls, name=name, config={})
    keras_json = dict(class_name='Model', keras_version='2.2.4', backend='tensorflow', config=dict(name='doc_encoder', layers=[
        layer('InputLayer', 'input_1'), layer('Embedding', 'embedding_1'), layer('Dropout', 'dropout_1'), layer('Conv1D', 'conv1d_1'),
        layer('Lambda', 'lambda_1'), layer('Masking', 'masking_1'), layer('Dropout', 'dropout_2'),
        layer('SimpleAttentionMaskSupport', 'simple_attention_mask_support_1'), layer('Dense', 'dense_1')]))
    jp, pp = str(tmp_path / 'encoder.json'), str(tmp_path / 'encoder.pkl')
    with open(jp, 'w') as f:
        json.dump(json.dumps(keras_json), f)             # the reference dumps the json STRING (utils.py:75-77)
    with open(pp, 'wb') as f:
        pickle.dump(W, f, protocol=pickle.HIGHEST_PROTOCOL)
    m = utils.load_model((jp, pp))
    assert m.config['keras'] and m.config['weight_names'] == ['word_emb', 'conv_w', 'conv_b', 'att_w', 'att_b', 'dense_w', 'dense_b']
    P = m.params()
    assert P['att_w'].shape == (F,) and np.array_equal(P['att_w'], W[3][:, 0]) and np.array_equal(P['conv_w'], W[1])
    # a recurrent layer in the json is not a cnnatt doc encoder: refuse
    bad = dict(keras_json, config=dict(name='x', layers=keras_json['config']['layers'] + [layer('GRU', 'gru_1')]))
    with open(jp, 'w') as f:
        json.dump(json.dumps(bad), f)
    with pytest.raises(ValueError):
        utils.load_model((jp, pp))
    # count mismatch
    with open(jp, 'w') as f:
        json.dump(json.dumps(keras_json), f)
    with open(pp, 'wb') as f:
        pickle.dump(W[:-1], f)
    with pytest.raises(ValueError):
        utils.load_model((jp, pp))


def test_vertsup_and_vertalt_generators():
    """Seq2VecPaperSoftmaxDaysIdVertSup samples carry the one-hot verticals of the W history slots + 1+K candidates
    (task/paper.py:897-902); ...VertAlt splits the documents 10 % / 90 %, serves (title, one-hot vertical) batches and
    multiplies config.epochs by config.round (:1003-1090)."""
    sh, d, emb, tok, cfg = _dataset()
    base = {k: getattr(cfg, k) for k in ('input_training_data_path', 'title_shape', 'window_size', 'negative_samples',
                                         'batch_size', 'textual_embedding_dim', 'title_filter_shape',
                                         'user_embedding_dim', 'debug')}
    h = task.get(settings.Config(dict(base, task='Seq2VecPaperSoftmaxDaysIdVertSup', arch='igru', days=100000)))
    s = next(h.train_gen())
    C = 1 + sh.K
    assert len(s) == 2 + C + 2 and s[-1].shape == (sh.W + C, len(utils.verticals)) and s[-2] == [1] + [0] * sh.K
    ids = s[-1].argmax(-1)
    assert np.all(s[-1].sum(-1) == 1) and np.all(ids[:sh.W][~s[1].any(-1)] == 0)          # pad slots carry 'N/A' = 0
    # the candidate verticals are those of the candidate titles
    vert_of_title = {tuple(doc.title.astype(int)): doc.vertical for doc in h.docs.values()}
    assert [vert_of_title[tuple(t.astype(int))] for t in s[2:2 + C]] == ids[sh.W:].tolist()
    (x, y) = next(h.valid)
    assert len(x) == 2 + C and len(y) == 2 and y[1].shape == (4, sh.W + C, len(utils.verticals))
    assert h.get_vertical_classifier() == (len(utils.verticals), h.config.hidden_dim)

    np.random.seed(5)
    a = task.get(settings.Config(dict(base, task='Seq2VecPaperSoftmaxDaysIdVertAlt', arch='igru', days=100000, round=4,
                                      epochs=3)))
    assert a.round == 4 and a.config.epochs == 12
    n = len(a.data_titles)
    assert n == sh.n_news and len(a.train_index) == n // 10 and len(a.valid_index) == n - n // 10
    assert sorted(np.concatenate([a.train_index, a.valid_index]).tolist()) == list(range(n))
    assert a.data_verticals.shape == (n, len(a.verticals)) and np.all(a.data_verticals.sum(-1) == 1)
    a.train_seq = False
    assert a.training_step == len(a.train_index) // 4 and a.validation_step == len(a.valid_index) // 4
    t, v = next(a.train_vert)
    assert t.shape == (4, sh.L) and v.shape == (4, len(a.verticals))
    gen = a.train
    assert next(gen)[0].shape == (4, sh.L)
    a.train_seq = True
    xb, yb = next(gen)
    assert len(xb) == 2 + C and yb.shape == (4, C) and a.training_step == a.config.training_step
    a.training_step = 5                                   # the base class assigns these in _load_data: ignored (:1059-1065)
    assert a.training_step == a.config.training_step
