"""CPU tests that pin the oracle (oracle/) — the reference has no tests or golden vectors for this path
(SURVEY.md §4, §8c), so the oracle is validated by: two independent implementations, hand-computable micro
cases, finite differences, structural invariants and committed golden fixtures."""
import math
import os

import numpy as np
import pytest
import torch

from mnexp_b200 import synth
from oracle import lstur_numpy as on
from oracle import lstur_torch as ot

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lstur_golden.npz'))
SH = synth.SHAPES['tiny']
TOK = GOLD['doc_tokens']


def case(arch, seed=4242, bseed=99):
    P = synth.make_weights(SH, arch=arch, bias_noise=0.05, seed=seed)
    (b,), _ = synth.make_batches(SH, 1, seed=bseed)
    return P, b, TOK[b['hist_doc']], TOK[b['cand_doc']]


@pytest.mark.parametrize('arch', ['igru', 'gru', 'hgru', 'nigru'])
def test_oracle_reproduces_golden(arch):
    assert np.array_equal(synth.make_docs(SH.n_news, SH.L, SH.vocab)[0], TOK)
    P, b, ct, cd = case(arch)
    for k in ('user', 'hist_doc', 'cand_doc'):
        assert np.array_equal(b[k], GOLD['%s/batch/%s' % (arch, k)])
    r = on.lstur_forward(P, b['user'], ct, cd, arch=arch, aux=True)
    for k in ('probs', 'logits', 'sigmoid', 'user_vec', 'cand_vec', 'hist_vec'):
        np.testing.assert_allclose(r[k], GOLD['%s/%s' % (arch, k)], rtol=1e-12, atol=1e-14)
    ora = ot.LsturOracle(P, arch=arch)
    loss, grads = ora.loss_and_grads(b['user'], ct, cd)
    assert abs(float(loss) - float(GOLD['%s/loss' % arch])) < 1e-12
    for k, g in grads.items():
        np.testing.assert_allclose(g.numpy(), GOLD['%s/grad/%s' % (arch, k)], rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize('arch', ['igru', 'gru', 'hgru', 'nigru', 'pgru', 'vo'])
def test_numpy_and_torch_implementations_agree(arch):
    P, b, ct, cd = case(arch, seed=7)
    pn = on.lstur_forward(P, b['user'], ct, cd, arch=arch)
    pt = ot.LsturOracle(P, arch=arch).forward(b['user'], ct, cd).detach().numpy()
    assert np.abs(pn - pt).max() < 1e-12
    p32 = ot.LsturOracle(P, arch=arch, dtype=torch.float32).forward(b['user'], ct, cd).detach().numpy()
    assert np.abs(pn - p32).max() < 1e-4 * np.abs(pn).max() + 1e-6


def test_scalar_loop_restatements():
    P, b, ct, cd = case('igru')
    P64 = {k: v.astype(np.float64) for k, v in P.items()}
    assert np.abs(on.news_encoder(TOK[:6], P) - on.news_encoder_loops(TOK[:6], P64)).max() < 1e-12
    g = np.random.default_rng(0)
    H = g.standard_normal((3, 4, SH.U))
    H[0, :2] = 0
    h0 = g.standard_normal((3, SH.U))
    a = on.gru_last_state(H, h0, P64['gru_wx'], P64['gru_wh'], P64['gru_b'])
    assert np.abs(a - on.gru_loops(H, h0, P64['gru_wx'], P64['gru_wh'], P64['gru_b'])).max() < 1e-12


def test_hand_computed_news_encoder():
    """L=3, E=1, F=1, k=3, all weights 1: C_t = x_{t-1}+x_t+x_{t+1} (zero outside), a=tanh(C), softmax-like pooling."""
    P = dict(word_emb=np.array([[0.], [1.], [2.]]), conv_w=np.ones((3, 1, 1)), conv_b=np.zeros(1), att_w=np.ones(1),
             att_b=np.zeros(1), dense_w=np.ones((1, 1)), dense_b=np.zeros(1))
    tok = np.array([[1, 2, 0]])                  # x = 1, 2, (pad row 0 = 0)
    C = np.array([1 + 2, 1 + 2 + 0, 0.0])        # position 2 is a pad token -> masked
    e = np.exp(np.tanh(C)) * np.array([1, 1, 0])
    want = (e / (e.sum() + 1e-7) * C).sum()
    assert abs(on.news_encoder(tok, P)[0, 0] - want) < 1e-12
    # mask_zero=False: a real token next to a pad convolves with row 0 of the table, whatever it holds
    P2 = dict(P, word_emb=np.array([[5.], [1.], [2.]]))
    C2 = np.array([1 + 2, 1 + 2 + 5, 0.0])
    e2 = np.exp(np.tanh(C2)) * np.array([1, 1, 0])
    assert abs(on.news_encoder(tok, P2)[0, 0] - (e2 / (e2.sum() + 1e-7) * C2).sum()) < 1e-12
    # all-pad title: pooled = 0, doc vector = dense bias
    assert on.news_encoder(np.zeros((1, 3), int), dict(P, dense_b=np.array([0.25])))[0, 0] == 0.25


def test_hand_computed_gru_step():
    """U=1, one step, all weights 1, h0=0.5: z=r=hs(x+h), hh=tanh(x+r*h), h'=z*h+(1-z)*hh."""
    hs = lambda v: min(1.0, max(0.0, 0.2 * v + 0.5))
    x, h = 2.0, 0.5
    z = r = hs(x + h)
    want = z * h + (1 - z) * math.tanh(x + r * h)
    got = on.gru_last_state(np.array([[[x]]]), np.array([[h]]), np.ones((1, 3)), np.ones((1, 3)), np.zeros(3))
    assert abs(got[0, 0] - want) < 1e-15
    # sigmoid recurrent activation (Keras >= 2.3) is a parameter
    sg = lambda v: 1 / (1 + math.exp(-v))
    z = r = sg(x + h)
    got = on.gru_last_state(np.array([[[x]]]), np.array([[h]]), np.ones((1, 3)), np.ones((1, 3)), np.zeros(3), 'sigmoid')
    assert abs(got[0, 0] - (z * h + (1 - z) * math.tanh(x + r * h))) < 1e-15


@pytest.mark.parametrize('arch', ['igru', 'gru'])
def test_autograd_matches_finite_differences(arch):
    P, b, ct, cd = case(arch, seed=11)
    ora = ot.LsturOracle(P, arch=arch)
    _, grads = ora.loss_and_grads(b['user'], ct, cd)
    g = np.random.default_rng(3)
    P64 = {k: v.astype(np.float64) for k, v in P.items()}
    for name in ('conv_w', 'att_w', 'dense_w', 'gru_wx', 'gru_wh', 'gru_b', 'user_emb'):
        d = g.standard_normal(P64[name].shape)
        if name == 'user_emb':
            d[np.setdiff1d(np.arange(d.shape[0]), b['user'])] = 0
        eps = 1e-6
        lp = on.lstur_loss(dict(P64, **{name: P64[name] + eps * d}), b['user'], ct, cd, arch=arch)
        lm = on.lstur_loss(dict(P64, **{name: P64[name] - eps * d}), b['user'], ct, cd, arch=arch)
        fd = (lp - lm) / (2 * eps)
        an = float((grads[name].numpy() * d).sum())
        assert abs(fd - an) < 1e-6 * max(1.0, abs(an)), name


def test_invariants():
    P, b, ct, cd = case('igru', seed=13)
    base = on.lstur_forward(P, b['user'], ct, cd, aux=True)
    # extra left padding of the history changes nothing (masked steps carry the state)
    ct2 = np.concatenate([np.zeros_like(ct[:, :3]), ct], 1)
    assert np.abs(on.lstur_forward(P, b['user'], ct2, cd) - base['probs']).max() < 1e-13
    # permuting the candidates permutes the probabilities
    perm = np.array([2, 0, 1])
    assert np.abs(on.lstur_forward(P, b['user'], ct, cd[:, perm]) - base['probs'][:, perm]).max() < 1e-13
    # all-masked history: ini -> h_T = h0 ; con (nigru) -> 0
    z = np.zeros_like(ct)
    assert np.array_equal(on.lstur_forward(P, b['user'], z, cd, aux=True)['user_vec'], P['user_emb'][b['user']].astype(np.float64))
    Pn = synth.make_weights(SH, arch='nigru', seed=13)
    assert np.all(on.lstur_forward(Pn, b['user'], z, cd, arch='nigru', aux=True)['user_vec'] == 0)
    # decomposed pipeline == full model (task/test_pipeline.py:214-265 `test_correct`)
    dv = on.news_encoder(ct.reshape(-1, SH.L), P).reshape(ct.shape[0], SH.W, -1) * on.history_mask(ct)[..., None]
    u = on.user_encoder('igru', b['user'], dv, P)
    s = on.score(u, on.news_encoder(cd.reshape(-1, SH.L), P).reshape(cd.shape[0], cd.shape[1], -1))
    assert np.abs(on.sigmoid(s) - base['sigmoid']).max() < 1e-13


def test_keras_adam_semantics():
    # first step: dp = -lr * g * sqrt(1-b2) / (sqrt(1-b2)|g| + eps)  ~ -lr*sign(g)
    g = np.array([0.3, -2.0, 1e-3])
    p, m, v = on.adam_step(np.zeros(3), g, np.zeros(3), np.zeros(3), 1, 1e-3)
    want = -1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-7)
    np.testing.assert_allclose(p, want, rtol=1e-12)
    assert np.abs(p + 1e-3 * np.sign(g)).max() < 5e-6          # |g| >> eps/sqrt(1-b2) = 3.2e-6
    # dense semantics: a row whose gradient is zero after one non-zero step keeps moving on its momentum
    p2, m2, v2 = on.adam_step(p, np.zeros(3), m, v, 2, 1e-3)
    assert np.all(np.abs(p2 - p) > 1e-5)
    # torch KerasAdam == numpy adam_step
    P = {'w': torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)}
    opt = ot.KerasAdam(P, lr=1e-3)
    opt.step({'w': torch.tensor(g)})
    np.testing.assert_allclose(P['w'].numpy(), np.array([1.0, 2.0, 3.0]) + p, rtol=1e-12)


def test_categorical_crossentropy_clip():
    y = np.array([[1.0, 0, 0]])
    assert abs(on.categorical_crossentropy(y, np.array([[0.5, 0.25, 0.25]])) - math.log(2)) < 1e-12
    assert abs(on.categorical_crossentropy(y, np.array([[0.0, 0.5, 0.5]])) + math.log(1e-7)) < 1e-9      # clipped
    assert abs(on.categorical_crossentropy(y, np.array([[2.0, 1.0, 1.0]])) - math.log(2)) < 1e-12       # renormalised


@pytest.mark.parametrize('arch,score_model', [('igru', 'dnn'), ('igru', 'ddot'), ('ngru', 'dnn'), ('ngru', 'ddot'),
                                              ('iigru', 'dot'), ('niavg', 'dot'), ('niavg', 'dnn')])
def test_remaining_archs_and_scorers_two_implementations(arch, score_model):
    """'dnn' / 'ddot' scorers (task/paper.py:448-455), 'ngru' (2U user vector), 'iigru' (two user tables), 'niavg'
    (masked mean): the float64 numpy loops and the torch restatement are written independently and must agree."""
    P = synth.make_weights(SH, arch=arch, bias_noise=0.05, seed=21, score_model=score_model)
    (b,), _ = synth.make_batches(SH, 1, seed=22)
    ct, cd = TOK[b['hist_doc']], TOK[b['cand_doc']]
    pn = on.lstur_forward(P, b['user'], ct, cd, arch=arch, score_model=score_model)
    pt = ot.LsturOracle(P, arch=arch, score_model=score_model).forward(b['user'], ct, cd).detach().numpy()
    assert pn.shape == (SH.B, 1 + SH.K) and np.abs(pn - pt).max() < 1e-12
    assert np.abs(pn.sum(-1) - 1).max() < 1e-12


def test_niavg_is_the_masked_mean_by_hand():
    """models.GlobalAveragePoolingMaskSupport (models.py:433-435): sum over the window / (number of unmasked steps +
    1e-7); masked (all-zero) steps contribute nothing to either."""
    H = np.zeros((2, 4, 3))
    H[0, 1] = [1.0, 2.0, 3.0]
    H[0, 3] = [3.0, 2.0, 1.0]
    u = on.user_encoder('niavg', np.zeros(2, dtype=int), H, {})
    assert np.allclose(u[0], np.array([4.0, 4.0, 4.0]) / (2 + 1e-7)) and np.all(u[1] == 0)
    ut = ot.user_encoder('niavg', torch.zeros(2, dtype=torch.long), torch.tensor(H), {})
    assert np.abs(ut.numpy() - u).max() < 1e-15


def test_weighted_bce_by_hand_and_gradient():
    """Seq2Vec.loss (task/seq2vec.py:213-216): -0.5 (1+K) mean(y log(p+1e-8) gain + (1-y) log(1-p+1e-8) / K)."""
    y = torch.tensor([[1.0], [0.0], [0.0]], dtype=torch.float64)
    p = torch.tensor([[0.8], [0.3], [0.6]], dtype=torch.float64, requires_grad=True)
    K, gain = 4, 1.5
    want = -0.5 * (1 + K) * (math.log(0.8 + 1e-8) * gain + math.log(0.7 + 1e-8) / K + math.log(0.4 + 1e-8) / K) / 3
    l = ot.weighted_bce(y, p, gain=gain, negative_samples=K)
    assert abs(float(l.detach()) - want) < 1e-14
    l.backward()
    g = p.grad.numpy().reshape(-1)
    want_g = -0.5 * (1 + K) / 3 * np.array([gain / (0.8 + 1e-8), -1 / (K * (0.7 + 1e-8)), -1 / (K * (0.4 + 1e-8))])
    assert np.abs(g - want_g).max() < 1e-12
    # sigmoid head of the torch oracle: one candidate per row
    P = synth.make_weights(SH, arch='igru', bias_noise=0.05, seed=5, score_model='dnn')
    (b,), _ = synth.make_batches(SH, 1, seed=6)
    ora = ot.LsturOracle(P, arch='igru', score_model='dnn')
    u, c, d = ora._ints(b['user'], TOK[b['hist_doc']], TOK[b['cand_doc']][:, :1])
    out = ot.forward(ora.P, u, c, d, arch='igru', score_model='dnn', head='sigmoid')
    assert out.shape == (SH.B, 1) and bool(((out > 0) & (out < 1)).all())


def test_variant_golden_fixtures_reproduce():
    """tests/golden/lstur_golden_variants.npz (make_golden_variants.py): remaining archs, scorers and the sigmoid family."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('mgv', os.path.join(os.path.dirname(__file__), 'golden', 'make_golden_variants.py'))
    mgv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgv)
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lstur_golden_variants.npz'))
    for arch, score_model, head in mgv.CASES:
        key = '%s-%s-%s' % (arch, score_model, head)
        _, _, _, _, probs, loss, grads = mgv.run_case(arch, score_model, head, SH, TOK)
        assert np.array_equal(probs, gold[key + '/probs']) and loss == float(gold[key + '/loss'])
        for k, g in grads.items():
            if g is not None and k != 'word_emb':
                assert np.array_equal(g.numpy(), gold[key + '/grad/' + k]), (key, k)


COOK_NEW = [('iavg', 'dnn'), ('iatt', 'ddot'), ('att', 'dot'), ('ilstm', 'dnn'), ('inagru', 'dot'), ('atgru', 'dnn'),
            ('algru', 'dot')]


@pytest.mark.parametrize('arch,score_model', COOK_NEW)
def test_cook_user_encoders_two_implementations(arch, score_model):
    """Cook.get_user_encoder branches iavg / iatt / ilstm / inagru / atgru / algru (task/cook.py:155-193) and
    Seq2VecPaper 'att' (task/paper.py:206-208): numpy float64 loops vs the independent torch restatement."""
    P = synth.make_weights(SH, arch=arch, bias_noise=0.05, seed=31, score_model=score_model)
    (b,), _ = synth.make_batches(SH, 1, seed=32)
    ct, cd = TOK[b['hist_doc']], TOK[b['cand_doc']]
    pn = on.lstur_forward(P, b['user'], ct, cd, arch=arch, score_model=score_model)
    pt = ot.LsturOracle(P, arch=arch, score_model=score_model).forward(b['user'], ct, cd).detach().numpy()
    assert pn.shape == (SH.B, 1 + SH.K) and np.abs(pn - pt).max() < 1e-12


def test_cook_heads_by_hand():
    """SimpleAttentionMaskSupport over two steps, AlphaAdd and the LSTM step, computed by hand."""
    # SimpleAttentionMaskSupport over a two-step sequence [h ; u], kernel k, bias 0 (models.py:474-489); an all-zero step
    # is masked out
    h = np.array([[1.0, 0.0]])
    u = np.array([[0.0, 2.0]])
    k = np.array([0.5, -0.25])
    a = np.tanh(np.array([0.5, -0.5]))
    e = np.exp(a)
    want = (e[0] * h + e[1] * u) / (e.sum() + 1e-7)
    got = on.masked_attention(np.stack([h, u], 1), k, 0.0)
    assert np.abs(got - want).max() < 1e-15
    got0 = on.masked_attention(np.stack([h, 0 * u], 1), k, 0.0)
    assert np.abs(got0 - h * e[0] / (e[0] + 1e-7)).max() < 1e-15
    # cook 'atgru' as the reference WRITES it (task/cook.py:184-190): the 2U entries of [GRU ; id] are one-feature steps,
    # zero entries are masked, the output is one scalar: sum_i x_i exp(tanh(k x_i)) / (sum_live exp(tanh(k x_i)) + 1e-7)
    P = dict(user_emb=np.array([[0.0, 2.0]]), uatt_w=np.array([0.5]), uatt_b=np.array([0.0]), gru_wx=np.zeros((2, 6)),
             gru_wh=np.zeros((2, 6)), gru_b=np.zeros(6))
    out = on.user_encoder('atgru', np.zeros(1, dtype=int), np.zeros((1, 3, 2)), P)      # GRU state stays 0: only x = 2 is live
    e2 = np.exp(np.tanh(0.5 * 2.0))
    assert out.shape == (1, 1) and abs(out[0, 0] - 2.0 * e2 / (e2 + 1e-7)) < 1e-15
    # algru (models.py:551-552)
    P = dict(user_emb=np.array([[2.0, 4.0]]), alpha=np.array([0.25]), gru_wx=np.zeros((2, 6)), gru_wh=np.zeros((2, 6)),
             gru_b=np.zeros(6))
    H = np.zeros((1, 3, 2))          # all-masked history: the GRU returns its zero initial state
    out = on.user_encoder('algru', np.zeros(1, dtype=int), H, P)
    assert np.allclose(out, 0.75 * P['user_emb'])
    # one LSTM step from zero state with identity-like weights: c = i * g, h = o * tanh(c)
    G = 1
    Wx = np.array([[1.0, 0.0, 2.0, 3.0]])
    b = np.array([0.0, 1.0, 0.0, 0.0])
    x = np.array([[[0.5]]])
    hs = lambda v: np.clip(0.2 * v + 0.5, 0, 1)
    c = hs(0.5) * np.tanh(1.0)
    want_h = hs(1.5) * np.tanh(c)
    got_h = on.lstm_last_state(x, Wx, np.zeros((G, 4 * G)), b)
    assert abs(got_h[0, 0] - want_h) < 1e-15
    # a trailing masked (all-zero) step carries the state
    x2 = np.concatenate([x, np.zeros((1, 1, 1))], 1)
    assert abs(on.lstm_last_state(x2, Wx, np.zeros((G, 4 * G)), b)[0, 0] - want_h) < 1e-15
