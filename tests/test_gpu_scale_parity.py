"""Full-scale parity: the table sizes and batch shapes of BASELINE.json configs C2 / C3 / C5 in the precision the
benchmark runs (fp16_tc), checked against the float64 oracle on a sample of the batch rows (the oracle needs seconds
per 64 impressions).  Tolerance: north_star's 1e-3, norm-wise and element-wise (tests/tolerances.py)."""
import numpy as np
import pytest
import torch

from mnexp_b200 import synth
from oracle import lstur_numpy as on
from tolerances import assert_close

pytestmark = pytest.mark.gpu


def sampled_forward_check(lib, sh, B, n_sample, precision='fp16_tc', seed=0, training=False):
    from mnexp_b200.engine import LsturEngine
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch=sh.arch, bias_noise=0.02)
    (b,), _ = synth.make_batches(sh, 1, seed=1236 + seed, B=B)
    eng = LsturEngine(P, B, sh.W, 1 + sh.K, sh.L, arch=sh.arch, doc_tokens=tok, precision=precision, training=training)
    db = eng.to_device_batch(b)
    eng.forward(db, training=training)
    if training:            # the full-size backward must at least run and produce finite gradients
        eng.backward(db)
        torch.cuda.synchronize()
        g = eng.dense_grad
        assert bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
    torch.cuda.synchronize()
    rows = np.sort(np.random.default_rng(seed).choice(B, n_sample, replace=False))
    ref = on.lstur_forward(P, b['user'][rows], tok[b['hist_doc'][rows]], tok[b['cand_doc'][rows]], arch=sh.arch, aux=True)
    D = eng.D
    dv = eng.view('doc_vec').reshape(-1, D)
    hist = dv[:B * sh.W].reshape(B, sh.W, D)[rows].cpu().numpy()
    cand = dv[B * sh.W:].reshape(B, 1 + sh.K, D)[rows].cpu().numpy()
    # 640 000 elements: the element-wise form is asserted for all but a 1e-5 tail, which must stay within 1.5 x the bound
    assert_close(hist, ref['hist_vec'], 1e-3, 'history vectors', tail=1e-5, tail_excess=1.5)
    assert_close(cand, ref['cand_vec'], 1e-3, 'candidate vectors', tail=1e-5, tail_excess=1.5)
    assert_close(eng.view('user_vec').reshape(B, -1)[rows].cpu().numpy(), ref['user_vec'], 1e-3, 'user vectors')
    assert_close(eng.view('logits').reshape(B, -1)[rows].cpu().numpy(), ref['logits'], 1e-3, 'scores')
    assert_close(eng.view('probs').reshape(B, -1)[rows].cpu().numpy(), ref['probs'], 1e-3, 'probabilities')
    lr = eng.view('loss_rows')[rows].cpu().numpy()
    ref_l = -np.log(np.clip(ref['probs'][:, 0] / ref['probs'].sum(-1), 1e-7, 1 - 1e-7))
    assert_close(lr, ref_l, 1e-3, 'per-row loss')


def test_c3_scale_forward_sampled_rows(lib):
    """C3: 1M users / 130k news / 100k vocab, B=1024, LSTUR-ini — the configuration bench.py times."""
    sampled_forward_check(lib, synth.SHAPES['C3'], 1024, 64)


def test_c5_full_batch_sampled_rows(lib):
    """C5 as benchmarked: W=200, L=50, B=2048 per rank (419 840 titles, 21 M tokens per step; the index arithmetic of
    every kernel crosses 2^31 elements here), forward + backward, 8 sampled rows against the float64 oracle."""
    sampled_forward_check(lib, synth.SHAPES['C5'], 2048, 8, seed=2, training=True)


def test_c2_scale_forward_sampled_rows(lib):
    """C2: LSTUR-con ('gru': Dense([GRU ‖ user])) at 50k users / 50k news / 30k vocab, B=1024."""
    sampled_forward_check(lib, synth.SHAPES['C2'], 1024, 64, seed=1)


def test_c5_shape_full_width_tensor_core(lib):
    """C5 (W=200, L=50, variable-length masks) at full width E300 / F400 / U200 on the tensor-core path (64-row title
    slots), small batch: forward vs the float64 oracle to the 1e-3 bound, gradients vs autograd (conv bias shifted so
    that no ReLU gate can flip, as in test_engine_fp16_tc_forward_and_grads)."""
    from mnexp_b200.engine import LsturEngine
    from oracle import lstur_torch as ot
    from tolerances import rel
    sh = synth.Shape('c5w', 300, 2000, 5000, L=50, W=200, K=4, B=4, E=300, F=400, U=200)
    assert lib.lstur_tc_supported(sh.L, sh.E, sh.F, 3) == 1
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=77)
    P['conv_b'] = P['conv_b'] + np.float32(1.0)
    (b,), frac = synth.make_batches(sh, 1, seed=78)
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='igru', doc_tokens=tok, precision='fp16_tc')
    db = eng.to_device_batch(b)
    eng.forward(db, training=True, seed=1)
    eng.backward(db)
    torch.cuda.synchronize()
    ora = ot.LsturOracle(P, arch='igru')
    out = ora.forward(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], aux=True)
    nh = sh.B * sh.W
    dv = eng.view('doc_vec').reshape(-1, eng.D).cpu().numpy()
    assert_close(dv[:nh], out['hist_vec'].detach().numpy().reshape(nh, -1), 1e-3, 'history vectors')
    assert_close(dv[nh:], out['cand_vec'].detach().numpy().reshape(-1, eng.D), 1e-3, 'candidate vectors')
    assert_close(eng.view('user_vec').reshape(sh.B, -1).cpu().numpy(), out['user_vec'].detach().numpy(), 1e-3, 'user vectors')
    assert_close(eng.view('probs').reshape(sh.B, -1).cpu().numpy(), out['probs'].detach().numpy(), 1e-3, 'probabilities')
    loss, ref = ora.loss_and_grads(b['user'], tok[b['hist_doc']], tok[b['cand_doc']])
    assert abs(eng.loss() - float(loss)) < 1e-3 * max(1.0, abs(float(loss)))
    got = eng.get_grads_dict()
    for k, g in ref.items():
        if g is None:
            continue
        assert rel(got[k], g.numpy()) < 3e-2, k          # W = 200 recurrent steps in 16-bit operands: grows like sqrt(W)
