"""GPU parity of the remaining Cook.get_user_encoder branches (task/cook.py:155-193: iavg, iatt, ilstm, inagru, atgru,
algru) and of Seq2VecPaper's 'att' user encoder (task/paper.py:206-208): forward outputs and every gradient through the
C-ABI against the float64 oracle, in the fp32 verification precision and (forward) in the tensor-core precision."""
import ctypes

import numpy as np
import pytest
import torch

from mnexp_b200 import synth
from oracle import lstur_torch as ot
from tolerances import rel, elem_excess

pytestmark = pytest.mark.gpu

CASES = [
    # engine arch, flavour, oracle arch, scorer
    ('iavg', 'cook', 'iavg', 'dnn'), ('iatt', 'cook', 'iatt', 'ddot'), ('ilstm', 'cook', 'ilstm', 'dnn'),
    ('inagru', 'cook', 'inagru', 'dot'), ('atgru', 'cook', 'atgru', 'dnn'), ('algru', 'cook', 'algru', 'dot'),
    ('att', 'sigmoid', 'att', 'dot'),
]


def _case(arch, oarch, score_model, seed=1301):
    sh = synth.SHAPES['tiny']
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch=oarch, bias_noise=0.05, seed=seed, score_model=score_model)
    (b,), _ = synth.make_batches(sh, 1, seed=seed + 1)
    return sh, tok, P, dict(b)


@pytest.mark.parametrize('arch,flavour,oarch,score_model', CASES)
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cook_user_encoders_match_oracle(lib, arch, flavour, oarch, score_model, precision):
    from mnexp_b200.engine import LsturEngine
    sh, tok, P, b = _case(arch, oarch, score_model)
    scale = None
    if flavour == 'cook':       # id_keep: u0 * Dropout(1 - id_keep)(idx_mask) (task/cook.py:141-142) as an explicit multiplier
        scale = (np.random.default_rng(9).random(sh.B) < 0.7).astype(np.float32) / np.float32(0.7)
        b['user_scale'] = scale
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch=arch, flavour=flavour, doc_tokens=tok, score_model=score_model,
                      precision=precision)
    db = eng.to_device_batch(b)
    probs = eng.forward(db, training=True, seed=1).cpu().numpy().copy()
    eng.backward(db)
    torch.cuda.synchronize()
    got = eng.get_grads_dict()
    ora = ot.LsturOracle(P, arch=oarch, score_model=score_model)
    u, c, d = ora._ints(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    kw = dict(arch=oarch, score_model=score_model, flavour='cook' if flavour == 'cook' else 'paper',
              u0_scale=None if scale is None else torch.tensor(scale, dtype=torch.float64)[:, None])
    out = ot.forward(ora.P, u, c, d, aux=True, **kw)
    tol = 5e-5 if precision == 'fp32' else 1e-3
    uv = eng.view('user_vec').reshape(sh.B, -1).cpu().numpy()
    assert uv.shape == tuple(out['user_vec'].shape)
    assert rel(uv, out['user_vec'].detach().numpy()) < tol
    assert elem_excess(uv, out['user_vec'].detach().numpy(), tol) <= 1.0
    assert rel(eng.view('logits').reshape(sh.B, -1).cpu().numpy(), out['logits'].detach().numpy()) < tol
    assert rel(probs, out['probs'].detach().numpy()) < tol
    loss = ot.loss_fn(ora.P, u, c, d, **kw)
    ref = dict(zip(ora.trainable, torch.autograd.grad(loss, [ora.P[k] for k in ora.trainable], allow_unused=True)))
    gtol = 5e-5 if precision == 'fp32' else 2e-2
    checked = 0
    gmax = max(float(g.abs().max()) for g in ref.values() if g is not None)
    for k, g in ref.items():
        if g is None or k in ('att_b', 'so_b'):
            continue
        assert k in got, k
        if k == 'uatt_b':      # sum of d a over the steps: cancels like att_b (softmax constraint); scale of its sibling uatt_w
            assert abs(float(np.asarray(got[k]).reshape(-1)[0]) - float(g.reshape(-1)[0])) < gtol * float(ref['uatt_w'].abs().max())
            continue
        if float(g.abs().max()) < 1e-9 * gmax:     # analytically zero (e.g. cook's linear 'ddot' biases under the softmax)
            assert float(np.abs(got[k]).max()) < 1e-6 * gmax, k
            continue
        assert rel(got[k], g.numpy()) < gtol, k
        checked += 1
    for k in {'iavg': ['user_emb'], 'iatt': ['uatt_w', 'uatt_b', 'user_emb'], 'ilstm': ['lstm_wx', 'lstm_wh', 'lstm_b'],
              'inagru': ['user_emb', 'user_emb2', 'gru_wh'], 'atgru': ['uatt_w', 'uatt_b', 'gru_wh', 'user_emb'],
              'algru': ['alpha', 'gru_wx', 'user_emb'], 'att': ['uatt_w', 'uatt_b']}[oarch]:
        assert ref[k] is not None and float(ref[k].abs().max()) > 0, k     # the head's own tensors carry a gradient


def test_alpha_add_constraint_after_update(lib):
    """AlphaAdd.alpha carries keras.constraints.MinMaxNorm(0, 1) (models.py:545): applied after the Adam update,
    alpha <- alpha * clip(|alpha|, 0, 1) / (1e-7 + |alpha|).  One training step vs the oracle's Keras-Adam + constraint."""
    from mnexp_b200.engine import LsturEngine
    sh, tok, P, b = _case('algru', 'algru', 'dot')
    P = dict(P, alpha=np.array([0.9995], dtype=np.float32))      # one lr=1e-3 step can push it over 1
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='algru', flavour='cook', doc_tokens=tok, lr=1e-3)
    db = eng.to_device_batch(b)
    ora = ot.LsturOracle(P, arch='algru', lr=1e-3)
    for _ in range(3):
        eng.train_step(db)
        ora.train_step(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], training=False)
        with torch.no_grad():
            a = ora.P['alpha']
            a.mul_(a.abs().clamp(0.0, 1.0) / (1e-7 + a.abs()))
    torch.cuda.synchronize()
    got = float(eng.get_weights_dict()['alpha'].reshape(-1)[0])
    want = float(ora.P['alpha'].reshape(-1)[0])
    assert got <= 1.0 and abs(got - want) < 1e-6, (got, want)


def test_lstm_kernels_match_autograd(lib):
    """lstur_lstm_fwd / lstur_lstm_bwd alone (ragged masks incl. fully masked rows and tiles) vs torch float64."""
    L = lib
    g = np.random.default_rng(3)
    B, W, G = 21, 9, 24
    XW = g.standard_normal((B, W, 4 * G)).astype(np.float32)
    Wh = (g.standard_normal((G, 4 * G)) * 0.3).astype(np.float32)
    gm = (g.random((B, W)) < 0.6).astype(np.float32)
    gm[3] = 0
    gm[8:16, :4] = 0
    dh = g.standard_normal((B, G)).astype(np.float32)
    t = lambda a: torch.as_tensor(a).cuda()
    dXW, dWh, dgm, ddh = t(XW), t(Wh), t(gm), t(dh)
    hT = torch.empty(B, G, device='cuda')
    S = [torch.empty(B, W, G, device='cuda') for _ in range(7)]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    assert L.lstur_lstm_fwd(B, W, G, p(dXW), p(dgm), p(dWh), 0, p(hT), G, *[p(x) for x in S], st) == 0
    WhT = dWh.t().contiguous()
    dA = torch.empty(B, W, 4 * G, device='cuda')
    SI, SF, SG, SO, SCP, SHP, STC = S
    assert L.lstur_lstm_bwd(B, W, G, p(dgm), p(SI), p(SF), p(SG), p(SO), p(SCP), p(STC), p(WhT), 0, p(ddh), G, p(dA), st) == 0
    torch.cuda.synchronize()
    x = torch.tensor(XW, dtype=torch.float64, requires_grad=True)
    w = torch.tensor(Wh, dtype=torch.float64, requires_grad=True)
    m = torch.tensor(gm) != 0
    h, c = torch.zeros(B, G, dtype=torch.float64), torch.zeros(B, G, dtype=torch.float64)
    for s in range(W):
        a = x[:, s] + h @ w
        i, f, gg, o = ot.hard_sigmoid(a[:, :G]), ot.hard_sigmoid(a[:, G:2 * G]), torch.tanh(a[:, 2 * G:3 * G]), ot.hard_sigmoid(a[:, 3 * G:])
        cn = f * c + i * gg
        hn = o * torch.tanh(cn)
        c = torch.where(m[:, s:s + 1], cn, c)
        h = torch.where(m[:, s:s + 1], hn, h)
    assert rel(hT.cpu().numpy(), h.detach().numpy()) < 2e-6
    (h * torch.tensor(dh, dtype=torch.float64)).sum().backward()
    assert rel(dA.cpu().numpy(), x.grad.numpy()) < 2e-5
    dWh_got = (SHP.reshape(-1, G).double().t() @ dA.reshape(-1, 4 * G).double()).cpu().numpy()
    assert rel(dWh_got, w.grad.numpy()) < 2e-5


# ---------------------------------------------------------------- Seq2VecPaperSoftmaxDaysIdVertSup / VertAlt (task/paper.py:884-1136)
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_vertsup_auxiliary_vertical_loss(lib, precision):
    """...VertSup: TimeDistributed Dense(hidden, relu) -> Dense(n_vert, softmax) over [history-masked clicked vectors ;
    candidate vectors], loss = CE_click + gain * mean CE_vertical (task/paper.py:954-990): both losses, the vertical
    probabilities and every gradient vs the float64 oracle; labels from the doc_vert table or per slot."""
    from mnexp_b200.engine import LsturEngine
    sh = synth.SHAPES['tiny']
    tok, vert, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    nv, hd, gain = 16, 10, 0.7
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=1401, vertsup=(nv, hd))
    (b,), _ = synth.make_batches(sh, 1, seed=1402)
    b = dict(b)
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='igru', doc_tokens=tok, doc_vert=vert, precision=precision,
                      aux_gain=gain)
    db = eng.to_device_batch(b)
    eng.forward(db, training=True, seed=1)
    eng.backward(db)
    torch.cuda.synchronize()
    ora = ot.LsturOracle(P, arch='igru')
    u, c, d = ora._ints(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    vl = (vert[b['hist_doc']], vert[b['cand_doc']])
    main, aux, vp = ot.loss_fn(ora.P, u, c, d, arch='igru', vert_labels=vl, aux_gain=gain, parts=True)
    tol = 5e-5 if precision == 'fp32' else 1e-3
    assert abs(eng.loss() - float(main)) <= tol * max(1.0, abs(float(main)))
    assert abs(eng.aux_loss() - float(aux)) <= tol * max(1.0, abs(float(aux)))
    n = sh.B * (sh.W + 1 + sh.K)
    got_vp = eng.view('vs_probs').reshape(n, nv).cpu().numpy()
    want_vp = torch.cat([vp[:, :sh.W].reshape(-1, nv), vp[:, sh.W:].reshape(-1, nv)]).detach().numpy()   # history rows first
    assert rel(got_vp, want_vp) < tol
    total = main + gain * aux
    ref = dict(zip(ora.trainable, torch.autograd.grad(total, [ora.P[k] for k in ora.trainable], allow_unused=True)))
    got = eng.get_grads_dict()
    gtol = 5e-5 if precision == 'fp32' else 2e-2
    for k, g in ref.items():
        if g is None or k == 'att_b':
            continue
        assert rel(got[k], g.numpy()) < gtol, k
    for k in ('vs_w1', 'vs_b1', 'vs_w2', 'vs_b2'):
        assert float(ref[k].abs().max()) > 0
    # per-slot labels (the generator protocol) give the same result as the doc_vert table
    eng2 = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='igru', doc_tokens=tok, precision=precision, aux_gain=gain)
    b2 = dict(b, hist_vert=vl[0], cand_vert=vl[1])
    db2 = eng2.to_device_batch(b2)
    eng2.forward(db2, training=True, seed=1)
    eng2.backward(db2)
    torch.cuda.synchronize()
    assert eng2.aux_loss() == eng.aux_loss()
    g2 = eng2.get_grads_dict()
    assert all(np.array_equal(g2[k], got[k]) for k in got)
    # without labels (test_model) the head is skipped and the click head is unchanged
    eng3 = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='igru', doc_tokens=tok, precision=precision, training=False)
    p3 = eng3.forward(eng3.to_device_batch(b), training=False).cpu().numpy()
    assert np.array_equal(p3, eng.view('probs').reshape(sh.B, -1).cpu().numpy())


@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_vertalt_title_classifier(lib, precision):
    """...VertAlt's vertical model: Dense(n_vert, softmax)(doc_encoder(title)) trained with its own Adam in alternation with
    the click model (task/paper.py:1128-1136).  Gradients vs autograd, then two alternating updates vs the oracle with two
    Keras-Adam instances over the shared weights."""
    from mnexp_b200.engine import LsturEngine
    sh = synth.SHAPES['tiny']
    tok, vert, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    nv = 16
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=1411, vertalt=nv)
    (b,), _ = synth.make_batches(sh, 1, seed=1412)
    ids = np.arange(1, 1 + 2 * sh.B)
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='igru', doc_tokens=tok, precision=precision, lr=1e-3)
    probs = eng.title_cls_forward(tok[ids], vert[ids], training=True, seed=3).cpu().numpy().copy()
    eng.title_cls_backward()
    torch.cuda.synchronize()
    ora = ot.LsturOracle(P, arch='igru', lr=1e-3)
    loss, vp = ot.title_cls_loss(ora.P, tok[ids], vert[ids], parts=True)
    tol = 5e-5 if precision == 'fp32' else 1e-3
    assert rel(probs, vp.detach().numpy()) < tol
    assert abs(float(eng.view('vc_loss')[0]) - float(loss)) <= tol * max(1.0, abs(float(loss)))
    names = ['conv_w', 'conv_b', 'att_w', 'dense_w', 'dense_b', 'vcls_w', 'vcls_b']
    ref = dict(zip(names, torch.autograd.grad(loss, [ora.P[k] for k in names])))
    got = eng.get_dense_grads_dict()          # the vertical model has no user side
    gtol = 5e-5 if precision == 'fp32' else 2e-2
    for k in names:
        assert rel(got[k], ref[k].numpy()) < gtol, k
    for k in ('gru_wx', 'gru_wh'):
        assert not got[k].any()
    if precision != 'fp32':
        return
    # alternate: click step, vertical step, click step — two optimizers over the shared doc encoder
    opt2 = ot.KerasAdam({k: ora.P[k] for k in names}, lr=1e-3)
    db = eng.to_device_batch(dict(b))
    for phase in ('seq', 'vert', 'seq', 'vert'):
        if phase == 'seq':
            eng.train_step(db)
            ora.train_step(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], training=False)
        else:
            eng.title_cls_train_step(tok[ids], vert[ids])
            l = ot.title_cls_loss(ora.P, tok[ids], vert[ids])
            opt2.step(dict(zip(names, torch.autograd.grad(l, [ora.P[k] for k in names]))))
    torch.cuda.synchronize()
    w = eng.get_weights_dict()
    for k in ('conv_w', 'dense_w', 'vcls_w', 'gru_wh', 'att_w'):
        assert np.abs(w[k] - ora.P[k].detach().numpy()).max() < 2e-5, k
