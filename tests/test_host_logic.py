"""Host-side logic (no GPU): reference data formats, Window / Impression semantics, generators, RNG replica."""
import os
import tempfile

import numpy as np

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
