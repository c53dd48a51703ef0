"""GPU parity tests: the CUDA path through the C-ABI vs the CPU oracle (oracle/).

Tolerances (BASELINE.json north_star): bit-exact for token gather and index dedup;
<= 1e-3 relative for encoder outputs, scores and loss.  The fp32 verification
precision is held to a much tighter 2e-5; relative error is max|a-b| / max|b|.
"""
import ctypes

import numpy as np
import pytest
import torch

from mnexp_b200 import rng, synth
from oracle import lstur_numpy as on
from oracle import lstur_torch as ot

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-5     # fp32 FFMA path vs float64 oracle
TOL_SPEC = 1e-3     # north_star tolerance


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def P_(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.fixture(scope='module')
def L(lib):
    assert torch.cuda.is_available()
    return lib


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------- integer paths (bit-exact)
@pytest.mark.parametrize('N,Lt', [(1, 30), (257, 30), (1000, 50), (33, 7)])
def test_token_gather_bit_exact(L, N, Lt):
    g = np.random.default_rng(N)
    docs = g.integers(0, 100000, (501, Lt)).astype(np.int32)
    docs[0] = 0
    ids = g.integers(0, 501, N).astype(np.int32)
    ids[0] = 0
    out = torch.empty((N, Lt), dtype=torch.int32, device='cuda')
    assert L.lstur_token_gather(N, Lt, 501, P_(dev(docs)), P_(dev(ids)), P_(out), stream()) == 0
    assert np.array_equal(out.cpu().numpy(), on.token_gather(docs, ids))


def test_token_gather_empty(L):
    assert L.lstur_token_gather(0, 30, 5, None, None, None, stream()) == 0


@pytest.mark.parametrize('n,hi', [(1, 5), (64, 10), (1000, 50), (1024, 1_000_000), (8192, 3000), (16384, 17)])
def test_sort_unique_bit_exact(L, n, hi):
    g = np.random.default_rng(n + hi)
    keys = g.integers(0, hi, n).astype(np.int32)
    k = dev(keys)
    sp = torch.empty(n, dtype=torch.int32, device='cuda')
    uq = torch.empty(n, dtype=torch.int32, device='cuda')
    ss = torch.empty(n + 1, dtype=torch.int32, device='cuda')
    inv = torch.empty(n, dtype=torch.int32, device='cuda')
    nu = torch.zeros(1, dtype=torch.int32, device='cuda')
    assert L.lstur_sort_unique_i32(n, P_(k), P_(sp), P_(uq), P_(ss), P_(inv), P_(nu), stream()) == 0
    u_ref, inv_ref, cnt_ref = np.unique(keys, return_inverse=True, return_counts=True)
    m = int(nu[0])
    assert m == len(u_ref)
    assert np.array_equal(uq[:m].cpu().numpy(), u_ref)
    assert np.array_equal(inv.cpu().numpy(), inv_ref)
    assert np.array_equal(np.diff(ss[:m + 1].cpu().numpy()), cnt_ref)
    assert np.array_equal(sp.cpu().numpy(), np.argsort(keys, kind='stable'))


def test_segment_sum_deterministic(L):
    g = np.random.default_rng(5)
    n, D = 1024, 200
    keys = g.integers(0, 100, n).astype(np.int32)
    src = g.standard_normal((n, D)).astype(np.float32)
    k, s = dev(keys), dev(src)
    sp = torch.empty(n, dtype=torch.int32, device='cuda'); uq = torch.empty_like(sp)
    ss = torch.empty(n + 1, dtype=torch.int32, device='cuda'); inv = torch.empty_like(sp)
    nu = torch.zeros(1, dtype=torch.int32, device='cuda')
    outs = []
    for _ in range(2):
        out = torch.zeros((n, D), device='cuda')
        assert L.lstur_sort_unique_i32(n, P_(k), P_(sp), P_(uq), P_(ss), P_(inv), P_(nu), stream()) == 0
        assert L.lstur_segment_sum_rows(n, D, P_(nu), P_(ss), P_(sp), P_(s), D, P_(out), stream()) == 0
        outs.append(out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    m = int(nu[0])
    ref = np.zeros((m, D))
    np.add.at(ref, np.unique(keys, return_inverse=True)[1], src.astype(np.float64))
    assert rel(outs[0][:m], ref) < 1e-6


# ---------------------------------------------------------------- GEMM building block
@pytest.mark.parametrize('ta,tb', [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (130, 70, 33), (900, 400, 5000), (257, 600, 200)])
def test_gemm_f32(L, ta, tb, M, N, K):
    g = np.random.default_rng(M * 7 + N * 3 + K + ta * 2 + tb)
    A = g.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    Bm = g.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = g.standard_normal(N).astype(np.float32)
    C = torch.zeros((M, N), device='cuda')
    nb = L.lstur_gemm_f32_workspace_bytes(M, N, K, None)
    ws = torch.empty(max(nb, 4), dtype=torch.uint8, device='cuda')
    assert L.lstur_gemm_f32(ta, tb, M, N, K, P_(dev(A)), A.shape[1], P_(dev(Bm)), Bm.shape[1], P_(C), N, P_(dev(bias)),
                            1, P_(ws), nb, stream()) == 0
    ref = np.maximum((A.T if ta else A).astype(np.float64) @ (Bm.T if tb else Bm).astype(np.float64) + bias, 0)
    assert rel(C.cpu().numpy(), ref) < 1e-5


# ---------------------------------------------------------------- whole path vs oracle
def make_case(shape_name='tiny', arch='igru', seed=0, B=None, **kw):
    sh = synth.SHAPES[shape_name]
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch=arch, bias_noise=0.05, seed=1237 + seed)
    (b,), _ = synth.make_batches(sh, 1, seed=1236 + seed, B=B)
    return sh, tok, P, b


def engine_for(sh, tok, P, arch, B=None, **kw):
    from mnexp_b200.engine import LsturEngine
    return LsturEngine(P, B or sh.B, sh.W, 1 + sh.K, sh.L, arch=arch, doc_tokens=tok, **kw)


@pytest.mark.parametrize('arch', ['igru', 'gru', 'hgru', 'nigru', 'pgru', 'vo'])
def test_forward_matches_oracle(L, arch):
    sh, tok, P, b = make_case(arch=arch)
    eng = engine_for(sh, tok, P, arch)
    db = eng.to_device_batch(b)
    probs = eng.forward(db, training=False).cpu().numpy()
    ref = on.lstur_forward(P, b['user'], tok[b['hist_doc']], tok[b['cand_doc']], arch=arch, aux=True)
    assert np.array_equal(eng.view('tokens', torch.int32).cpu().numpy().reshape(-1, sh.L),
                          np.concatenate([tok[b['hist_doc']].reshape(-1, sh.L), tok[b['cand_doc']].reshape(-1, sh.L)]))
    dv = eng.view('doc_vec').reshape(-1, eng.D).cpu().numpy()
    nh = sh.B * sh.W
    assert rel(dv[:nh], ref['hist_vec'].reshape(nh, -1)) < TOL_FP32
    assert rel(dv[nh:], ref['cand_vec'].reshape(-1, eng.D)) < TOL_FP32
    assert rel(eng.view('user_vec').reshape(sh.B, -1).cpu().numpy(), ref['user_vec']) < TOL_FP32
    assert rel(eng.view('logits').reshape(sh.B, -1).cpu().numpy(), ref['logits']) < TOL_FP32
    assert rel(probs, ref['probs']) < TOL_FP32
    loss_ref = on.categorical_crossentropy(np.eye(1 + sh.K)[np.zeros(sh.B, int)], ref['probs'])
    assert abs(eng.loss() - loss_ref) < TOL_FP32 * max(1, abs(loss_ref))


def test_forward_token_protocol_equals_docid_protocol(L):
    """Reference data protocol (clicked (B,W,L) float64 token arrays, task/paper.py:538-541) == doc-id protocol."""
    sh, tok, P, b = make_case()
    eng = engine_for(sh, tok, P, 'igru')
    p1 = eng.forward(eng.to_device_batch(b)).cpu().numpy().copy()
    b2 = dict(user=b['user'], hist_tok=tok[b['hist_doc']].astype(np.float64), cand_tok=tok[b['cand_doc']].astype(np.float64))
    p2 = eng.forward(eng.to_device_batch(b2)).cpu().numpy()
    assert np.array_equal(p1, p2)


@pytest.mark.parametrize('arch', ['igru', 'gru', 'hgru', 'nigru'])
def test_gradients_match_autograd(L, arch):
    sh, tok, P, b = make_case(arch=arch, seed=3)
    eng = engine_for(sh, tok, P, arch)
    db = eng.to_device_batch(b)
    eng.forward(db, training=True, seed=1)
    eng.backward(db)
    torch.cuda.synchronize()
    got = eng.get_grads_dict()
    ora = ot.LsturOracle(P, arch=arch)
    _, ref = ora.loss_and_grads(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    for k, g in ref.items():
        if g is None:
            continue
        assert k in got, k
        if k == 'att_b':
            # scalar sum of d a over every token: the softmax constraint makes its terms cancel almost exactly, so the
            # error is bounded against the scale of the un-cancelled sibling d att_w (same terms weighted by C)
            scale = max(abs(float(g)), float(np.abs(ref['att_w'].numpy()).max()))
            assert abs(float(np.asarray(got[k]).reshape(-1)[0]) - float(g)) < 5e-5 * scale, k
            continue
        assert rel(got[k], g.numpy()) < 5e-5, k


def test_all_padded_history_rows(L):
    """All-masked history => h_T = h0 (ini): SURVEY §8c invariant; also exercises the fully-masked GRU tile path."""
    sh, tok, P, b = make_case()
    b = dict(b)
    b['hist_doc'] = b['hist_doc'].copy()
    b['hist_doc'][:] = 0
    eng = engine_for(sh, tok, P, 'igru')
    eng.forward(eng.to_device_batch(b))
    uv = eng.view('user_vec').reshape(sh.B, -1).cpu().numpy()
    assert np.array_equal(uv, P['user_emb'][b['user']])


def test_train_steps_match_oracle_adam(L):
    sh, tok, P, _ = make_case(seed=5)
    batches, _ = synth.make_batches(sh, 3, seed=77)
    eng = engine_for(sh, tok, P, 'igru', lr=1e-3, sparse_user_adam=False)
    ora = ot.LsturOracle(P, arch='igru', lr=1e-3)
    for b in batches:
        lg = float(eng.train_step(eng.to_device_batch(b))[0])
        lo = ora.train_step(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], training=False)
        assert abs(lg - lo) < 1e-4 * max(1, abs(lo))
    w = eng.get_weights_dict()
    for k in ora.trainable:
        # Adam's first steps move every weight by ~lr regardless of gradient size, so compare on the update scale
        assert np.abs(w[k] - ora.P[k].detach().numpy()).max() < 2e-5, k


def test_sparse_user_adam_touches_only_batch_rows(L):
    sh, tok, P, b = make_case(seed=6)
    eng = engine_for(sh, tok, P, 'igru', sparse_user_adam=True)
    eng.train_step(eng.to_device_batch(b))
    w = eng.get_weights_dict()['user_emb']
    changed = np.where(np.abs(w - P['user_emb']).max(1) > 0)[0]
    assert set(changed) <= set(np.unique(b['user']))


def test_dropout_mask_replay(L):
    """Training-mode dropout (task/paper.py:147,158) with the device RNG replayed into the oracle."""
    sh, tok, P, b = make_case(seed=8)
    p = 0.2
    eng = engine_for(sh, tok, P, 'igru', dropout=p)
    db = eng.to_device_batch(b)
    seed = 11
    eng.forward(db, training=True, seed=seed)
    N = sh.B * (sh.W + 1 + sh.K)
    toks = np.concatenate([tok[b['hist_doc']].reshape(-1, sh.L), tok[b['cand_doc']].reshape(-1, sh.L)])
    dx = rng.dropout_multiplier(seed * 2, N * sh.L * sh.E, p).reshape(N, sh.L, sh.E)
    dc = rng.dropout_multiplier(seed * 2 + 1, N * sh.L * sh.F, p).reshape(N, sh.L, sh.F)
    ref = on.news_encoder(toks, P, drop_x=dx, drop_c=dc)
    dv = eng.view('doc_vec').reshape(N, -1).cpu().numpy()
    hm = (toks[:sh.B * sh.W] != 0).any(-1)
    ref[:sh.B * sh.W] *= hm[:, None]
    assert rel(dv, ref) < TOL_FP32
    frac = float((dx == 0).mean())
    assert abs(frac - p) < 0.02


def test_c1_shape_forward_and_grads(L):
    """BASELINE config C1 (LSTUR-ini, 1k users/5k docs, L30 E300 W50 K4, B64) — forward + grads vs oracle."""
    sh, tok, P, b = make_case('C1')
    eng = engine_for(sh, tok, P, 'igru')
    db = eng.to_device_batch(b)
    eng.forward(db, training=True)
    eng.backward(db)
    ora = ot.LsturOracle(P, arch='igru', dtype=torch.float64)
    out = ora.forward(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], aux=True)
    assert rel(eng.view('probs').reshape(sh.B, -1).cpu().numpy(), out['probs'].detach().numpy()) < TOL_FP32
    _, ref = ora.loss_and_grads(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    got = eng.get_grads_dict()
    for k, g in ref.items():
        assert rel(got[k], g.numpy()) < 1e-4, k


# ---------------------------------------------------------------- cook.py variant: vertical concat + id_keep
@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_cook_vertical_concat_and_id_keep(L, precision):
    """Cook.get_doc_encoder (task/cook.py:99-113): d' = [title_vec(F) | Vemb[vert] | Semb[subvert]], no Dense;
    history concat masked by the title tokens (:236-250); user vector scaled by Dropout(1-id_keep)(idx_mask)
    (:141-142, here an explicit per-row multiplier).  Forward vs the float64 oracle, gradients vs autograd."""
    import dataclasses
    from mnexp_b200.engine import LsturEngine
    dv, ds = 3, 5
    base = synth.SHAPES['tiny']
    sh = dataclasses.replace(base, name='cook', U=base.F + dv + ds)
    tok, vert, subvert = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, cook=True, dv=dv, ds=ds)
    (b,), _ = synth.make_batches(sh, 1, seed=77)
    g = np.random.default_rng(5)
    scale = (g.random(sh.B) < 0.7).astype(np.float32) / np.float32(0.7)          # idx_mask * dropout / id_keep
    b = dict(b, user_scale=scale)
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='ingru', flavour='cook', doc_tokens=tok, doc_vert=vert,
                      doc_subvert=subvert, precision=precision)
    assert eng.D == sh.F + dv + ds
    db = eng.to_device_batch(b)
    probs = eng.forward(db, training=True, seed=1).cpu().numpy().copy()          # dropout 0: training only saves state
    kw = dict(hist_vert=vert[b['hist_doc']], hist_subvert=subvert[b['hist_doc']], cand_vert=vert[b['cand_doc']],
              cand_subvert=subvert[b['cand_doc']])
    ref = on.lstur_forward(P, b['user'], tok[b['hist_doc']], tok[b['cand_doc']], arch='igru', aux=True,
                           u0_scale=scale[:, None].astype(np.float64), **kw)
    tol = TOL_FP32 if precision == 'fp32' else TOL_SPEC
    docv = eng.view('doc_vec').reshape(-1, eng.D).cpu().numpy()
    nh = sh.B * sh.W
    assert rel(docv[nh:], ref['cand_vec'].reshape(-1, eng.D)) < tol
    assert np.array_equal(docv[nh:, sh.F:sh.F + dv], P['vert_emb'][vert[b['cand_doc']].reshape(-1)])      # gathers are exact
    assert np.array_equal(docv[nh:, sh.F + dv:], P['subvert_emb'][subvert[b['cand_doc']].reshape(-1)])
    assert rel(docv[:nh], ref['hist_vec'].reshape(nh, -1)) < tol
    assert rel(probs, ref['probs']) < tol
    if precision != 'fp32':
        return
    eng.backward(db)
    torch.cuda.synchronize()
    got = eng.get_grads_dict()
    Pt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=(k != 'word_emb')) for k, v in P.items()}
    t = lambda x: torch.as_tensor(np.asarray(x)).long()
    loss = ot.loss_fn(Pt, t(b['user']), t(tok[b['hist_doc']]), t(tok[b['cand_doc']]), arch='igru',
                      u0_scale=torch.tensor(scale[:, None], dtype=torch.float64), **kw)
    names = [k for k in Pt if k != 'word_emb']
    grads = dict(zip(names, torch.autograd.grad(loss, [Pt[k] for k in names], allow_unused=True)))
    for k in ('vert_emb', 'subvert_emb', 'conv_w', 'gru_wx', 'gru_wh', 'user_emb', 'att_w'):
        assert rel(got[k], grads[k].numpy()) < 5e-5, k


# ---------------------------------------------------------------- remaining user encoders and scorers of §8 a8/a9
@pytest.mark.parametrize('arch,score_model,precision', [
    ('igru', 'dnn', 'fp32'), ('igru', 'ddot', 'fp32'), ('ngru', 'dnn', 'fp32'), ('ngru', 'ddot', 'fp32'),
    ('dgru', 'ddot', 'fp32'), ('niavg', 'dot', 'fp32'), ('niavg', 'dnn', 'fp32'), ('gru', 'ddot', 'fp16_tc'),
    ('ngru', 'dnn', 'fp16_tc'), ('iigru', 'dot', 'fp32'), ('iigru', 'dnn', 'fp16_tc'),
])
def test_scorers_and_remaining_archs_match_oracle(L, arch, score_model, precision):
    """'dnn' / 'ddot' scorers (task/paper.py:448-455), 'ngru' / 'dgru' (2U user vector, :600-611, only scorable by
    those two) and 'niavg' (:627-628): forward outputs, test head and every gradient against the float64 oracle."""
    sh = synth.SHAPES['tiny']
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch=arch, bias_noise=0.05, seed=1250, score_model=score_model)
    (b,), _ = synth.make_batches(sh, 1, seed=1249)
    b = dict(b)
    scale = None
    if arch == 'dgru':      # Dropout(0.5, noise_shape=(None, 1)) on the user vector: a per-sample multiplier in {0, 2}
        scale = (np.random.default_rng(5).random(sh.B) < 0.5).astype(np.float32) * 2.0
        b['user_scale'] = scale
    eng = engine_for(sh, tok, P, arch, score_model=score_model, precision=precision)
    db = eng.to_device_batch(b)
    probs = eng.forward(db, training=True, seed=1).cpu().numpy().copy()
    sig = eng.score_sigmoid().cpu().numpy()
    eng.backward(db)
    torch.cuda.synchronize()
    got = eng.get_grads_dict()
    ora = ot.LsturOracle(P, arch=arch, score_model=score_model)
    u, c, d = ora._ints(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    kw = dict(arch=arch, score_model=score_model, u0_scale=None if scale is None else torch.tensor(scale, dtype=torch.float64)[:, None])
    out = ot.forward(ora.P, u, c, d, aux=True, **kw)
    tol = 5e-5 if precision == 'fp32' else TOL_SPEC
    assert rel(eng.view('user_vec').reshape(sh.B, -1).cpu().numpy(), out['user_vec'].detach().numpy()) < tol
    assert rel(eng.view('logits').reshape(sh.B, -1).cpu().numpy(), out['logits'].detach().numpy()) < tol
    assert rel(probs, out['probs'].detach().numpy()) < tol
    assert rel(sig, torch.sigmoid(out['logits']).detach().numpy()) < tol
    loss = ot.loss_fn(ora.P, u, c, d, **kw)
    ref = dict(zip(ora.trainable, torch.autograd.grad(loss, [ora.P[k] for k in ora.trainable], allow_unused=True)))
    gtol = 5e-5 if precision == 'fp32' else 2e-2
    for k, g in ref.items():
        if g is None or k == 'att_b':
            continue
        assert k in got, k
        if k == 'so_b':     # the softmax is shift-invariant: d so_b = sum of d logits = 0 up to rounding
            assert abs(float(np.asarray(got[k]).reshape(-1)[0])) < 1e-6
            continue
        assert rel(got[k], g.numpy()) < gtol, k


# ---------------------------------------------------------------- sigmoid family: weighted BCE head (a11)
@pytest.mark.parametrize('arch,oarch,score_model,precision', [
    ('igru', 'igru', 'dot', 'fp32'), ('gru', 'ngru', 'dnn', 'fp32'), ('iigru', 'iicat', 'dnn', 'fp32'),
    ('nigru', 'nigru', 'dot', 'fp32'), ('niavg', 'niavg', 'dnn', 'fp32'), ('vo', 'vo', 'ddot', 'fp32'),
    ('igru', 'igru', 'dnn', 'fp16_tc'),
])
def test_sigmoid_family_weighted_bce(L, arch, oarch, score_model, precision):
    """Seq2VecPaper / Seq2VecPaperDot / Seq2VecPaperId (task/paper.py:222-383): one candidate per row, sigmoid head,
    Seq2Vec.loss (task/seq2vec.py:213-216) with gain and negative_samples; outputs, loss and gradients vs the oracle."""
    sh = synth.SHAPES['tiny']
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch=oarch, bias_noise=0.05, seed=1260, score_model=score_model)
    (b,), _ = synth.make_batches(sh, 1, seed=1261)
    g = np.random.default_rng(9)
    y = (g.random((sh.B, 1)) < 0.4).astype(np.float32)
    b = dict(user=b['user'], hist_doc=b['hist_doc'], cand_doc=b['cand_doc'][:, :1], label=y)
    from mnexp_b200.engine import LsturEngine
    eng = LsturEngine(P, sh.B, sh.W, 1, sh.L, arch=arch, flavour='sigmoid', doc_tokens=tok, score_model=score_model,
                      precision=precision, loss='bce', gain=1.7, bce_neg=sh.K)
    db = eng.to_device_batch(b)
    p_gpu = eng.forward(db, training=True, seed=1).cpu().numpy().copy()
    loss_gpu = eng.loss()
    eng.backward(db)
    torch.cuda.synchronize()
    got = eng.get_grads_dict()
    ora = ot.LsturOracle(P, arch=oarch, score_model=score_model)
    u, c, d = ora._ints(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    out = ot.forward(ora.P, u, c, d, arch=oarch, score_model=score_model, head='sigmoid')
    loss = ot.weighted_bce(torch.tensor(y, dtype=torch.float64), out, gain=1.7, negative_samples=sh.K)
    tol = 5e-5 if precision == 'fp32' else TOL_SPEC
    assert rel(p_gpu, out.detach().numpy()) < tol
    assert abs(loss_gpu - float(loss.detach())) < tol * max(1.0, abs(float(loss.detach())))
    ref = dict(zip(ora.trainable, torch.autograd.grad(loss, [ora.P[k] for k in ora.trainable], allow_unused=True)))
    gtol = 1e-4 if precision == 'fp32' else 2e-2
    for k, gr in ref.items():
        if gr is None or k == 'att_b':
            continue
        assert k in got, k
        assert rel(got[k], gr.numpy()) < gtol, k


def test_c5_shape_class_long_window_long_titles(L):
    """C5 of BASELINE.json (W=200, L=50, variable-length masks) at reduced width on the fp32 verification chain: the
    recurrence runs 200 steps.  (Full width on the tensor-core path: tests/test_gpu_scale_parity.py.)"""
    sh = synth.Shape('c5s', 60, 90, 150, L=50, W=200, K=4, B=6, E=16, F=32, U=16)
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=77)
    (b,), frac = synth.make_batches(sh, 1, seed=78)
    eng = engine_for(sh, tok, P, 'igru')
    db = eng.to_device_batch(b)
    probs = eng.forward(db, training=True, seed=1).cpu().numpy().copy()
    eng.backward(db)
    torch.cuda.synchronize()
    got = eng.get_grads_dict()
    ora = ot.LsturOracle(P, arch='igru')
    out = ora.forward(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], aux=True)
    assert rel(probs, out['probs'].detach().numpy()) < 5e-5
    assert rel(eng.view('user_vec').reshape(sh.B, -1).cpu().numpy(), out['user_vec'].detach().numpy()) < 5e-5
    _, ref = ora.loss_and_grads(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    for k, g in ref.items():
        if g is None or k == 'att_b':
            continue
        assert rel(got[k], g.numpy()) < 1e-4, k
