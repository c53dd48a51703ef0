"""GRU recurrence kernels (cluster / shared-memory-resident and streaming) vs the float64 oracle.

Reference: keras.layers.GRU(U)(Masking()(clicked), initial_state=user_vec) task/paper.py:596-613; semantics SURVEY §9.4.
The kernels take XW = H.Wx + b precomputed, so the oracle is driven with H = XW, Wx = I, b = 0 (masked steps = zero rows).
Tolerance: fp32 recurrence vs float64, 2e-5 relative (max|a-b| / max|b|)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import lstur_torch as ot

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def P_(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def make(B, W, G, seed, ini=True, pad='left'):
    g = np.random.default_rng(seed)
    XW = g.standard_normal((B, W, 3 * G)) * 0.8
    lens = g.integers(0, W + 1, B)
    lens[0] = W
    if B > 1:
        lens[1] = 0          # all-masked history: output = h0
    gm = np.zeros((B, W))
    for b in range(B):
        if pad == 'left':
            gm[b, W - lens[b]:] = 1
        else:                # arbitrary holes (Masking() is per step, not only left padding)
            gm[b] = g.random(W) < 0.6
    XW *= gm[:, :, None]
    Wh = g.standard_normal((G, 3 * G)) / np.sqrt(G)
    h0 = g.uniform(-0.5, 0.5, (B, G)) if ini else None
    dhT = g.standard_normal((B, G))
    return XW, gm, Wh, h0, dhT


def oracle(XW, Wh, h0, dhT, act):
    G = Wh.shape[0]
    H = torch.tensor(XW, dtype=torch.float64, requires_grad=True)
    Wh_t = torch.tensor(Wh, dtype=torch.float64)
    h0_t = torch.tensor(h0, dtype=torch.float64, requires_grad=True) if h0 is not None else None
    hT = ot.gru_last_state(H, h0_t, torch.eye(3 * G, dtype=torch.float64), Wh_t, torch.zeros(3 * G, dtype=torch.float64), act)
    (hT * torch.tensor(dhT)).sum().backward()
    return hT.detach().numpy(), H.grad.numpy(), (h0_t.grad.numpy() if h0 is not None else None)


def run(lib, impl, XW, gm, Wh, h0, dhT, act, order=None):
    B, W, G3 = XW.shape
    G = G3 // 3
    f = lambda a: torch.tensor(a, dtype=torch.float32).cuda().contiguous()
    XWd, gmd, Whd, dhTd = f(XW), f(gm), f(Wh), f(dhT)
    h0d = f(h0) if h0 is not None else None
    WhT = Whd.t().contiguous()
    hT = torch.full((B, G), float('nan'), device='cuda')
    sv = [torch.full((B, W, G), float('nan'), device='cuda') for _ in range(5)]
    dA = torch.full((B, W, 3 * G), float('nan'), device='cuda')
    dh0 = torch.full((B, G), float('nan'), device='cuda')
    a = 0 if act == 'hard_sigmoid' else 1
    od = torch.tensor(order, dtype=torch.int32).cuda() if order is not None else None
    if impl in ('tcb', 'tctc'):        # tensor-core backward after the fp32 cluster forward ('tcb') or the tensor-core one
        assert lib.lstur_gru_tc_supported(B, W, G) == 1
        fwd = lib.lstur_gru_fwd_cluster if impl == 'tcb' else lib.lstur_gru_fwd_tc
        rc = fwd(B, W, G, P_(XWd), P_(gmd), P_(h0d), G, P_(Whd), a, P_(hT), G, *[P_(s) for s in sv], P_(od), stream())
        assert rc == 0, lib.lstur_last_error()
        dbp = torch.full((lib.lstur_gru_tc_db_rows(B), 3 * G), float('nan'), device='cuda')
        rc = lib.lstur_gru_bwd_tc(B, W, G, P_(gmd), *[P_(s) for s in sv[:4]], P_(Whd), a, P_(dhTd), G, P_(dA), P_(dh0), G, P_(od), P_(dbp), stream())
        torch.cuda.synchronize()
        # the fused bias-gradient partials add up to the column sums of dA
        assert torch.allclose(dbp.sum(0), dA.reshape(-1, 3 * G).sum(0), rtol=1e-3, atol=1e-5 * float(dA.abs().max()) * B * W)
        assert rc == 0, lib.lstur_last_error()
    elif impl == 'tc':        # tensor-core forward; its saved tensors feed the cluster backward
        assert lib.lstur_gru_tc_supported(B, W, G) == 1
        rc = lib.lstur_gru_fwd_tc(B, W, G, P_(XWd), P_(gmd), P_(h0d), G, P_(Whd), a, P_(hT), G, *[P_(s) for s in sv], P_(od), stream())
        assert rc == 0, lib.lstur_last_error()
        rc = lib.lstur_gru_bwd_cluster(B, W, G, P_(gmd), *[P_(s) for s in sv[:4]], P_(WhT), a, P_(dhTd), G, P_(dA), P_(dh0), G, P_(od), stream())
        assert rc == 0, lib.lstur_last_error()
    elif impl == 'cluster':
        assert lib.lstur_gru_cluster_supported(B, W, G) == 1
        rc = lib.lstur_gru_fwd_cluster(B, W, G, P_(XWd), P_(gmd), P_(h0d), G, P_(Whd), a, P_(hT), G, *[P_(s) for s in sv], P_(od), stream())
        assert rc == 0, lib.lstur_last_error()
        rc = lib.lstur_gru_bwd_cluster(B, W, G, P_(gmd), *[P_(s) for s in sv[:4]], P_(WhT), a, P_(dhTd), G, P_(dA), P_(dh0), G, P_(od), stream())
        assert rc == 0, lib.lstur_last_error()
    else:
        rc = lib.lstur_gru_fwd_streaming(B, W, G, P_(XWd), P_(gmd), P_(h0d), G, P_(Whd), a, P_(hT), G, *[P_(s) for s in sv], stream())
        assert rc == 0, lib.lstur_last_error()
        rc = lib.lstur_gru_bwd_streaming(B, W, G, P_(gmd), *[P_(s) for s in sv[:4]], P_(WhT), a, P_(dhTd), G, P_(dA), P_(dh0), G, stream())
        assert rc == 0, lib.lstur_last_error()
    torch.cuda.synchronize()
    return hT.cpu().numpy(), dA.cpu().numpy(), dh0.cpu().numpy(), [s.cpu().numpy() for s in sv]


@pytest.mark.parametrize('impl', ['cluster', 'streaming'])
@pytest.mark.parametrize('B,W,G,ini,act,pad', [
    (1, 3, 8, True, 'hard_sigmoid', 'left'),
    (7, 5, 8, False, 'hard_sigmoid', 'holes'),
    (64, 50, 200, True, 'hard_sigmoid', 'left'),       # C1: cluster of 4, 50 units per CTA
    (50, 12, 100, False, 'sigmoid', 'holes'),          # hgru width (U/2): cluster of 2
    (300, 20, 64, True, 'hard_sigmoid', 'left'),       # one CTA per tile
    (1024, 50, 200, True, 'hard_sigmoid', 'left'),     # C3 batch: 32 rows per cluster, 8 rows per thread
    (130, 9, 256, True, 'sigmoid', 'left'),            # cluster of 8
])
def test_gru_recurrence_vs_oracle(lib, impl, B, W, G, ini, act, pad):
    XW, gm, Wh, h0, dhT = make(B, W, G, seed=B + W + G, ini=ini, pad=pad)
    hT_ref, dA_ref, dh0_ref = oracle(XW, Wh, h0, dhT, act)
    hT, dA, dh0, sv = run(lib, impl, XW, gm, Wh, h0, dhT, act)
    assert rel(hT, hT_ref) < TOL
    assert rel(dA, dA_ref) < 5 * TOL
    if ini:
        assert rel(dh0, dh0_ref) < 5 * TOL
    assert all(np.isfinite(s).all() for s in sv)          # the weight-gradient GEMMs read every saved row
    assert np.all(dA[gm == 0] == 0)


@pytest.mark.parametrize('B,W,G,ini,act,pad', [
    (9, 5, 11, True, 'hard_sigmoid', 'left'),          # U + vertical_embedding_dim of the tiny reference-run case
    (40, 12, 215, True, 'hard_sigmoid', 'holes'),      # ...DaysIdVert at the reference's defaults: 200 + 15 (task/paper.py:1204-1208)
    (17, 7, 33, False, 'sigmoid', 'left'),
])
def test_gru_streaming_any_width(lib, B, W, G, ini, act, pad):
    """widths that are no multiple of 4 run on the streaming kernels (padded shared-memory rows)"""
    assert lib.lstur_gru_cluster_supported(B, W, G) == 0 and lib.lstur_gru_tc_supported(B, W, G) == 0
    XW, gm, Wh, h0, dhT = make(B, W, G, seed=B + W + G, ini=ini, pad=pad)
    hT_ref, dA_ref, dh0_ref = oracle(XW, Wh, h0, dhT, act)
    hT, dA, dh0, sv = run(lib, 'streaming', XW, gm, Wh, h0, dhT, act)
    assert rel(hT, hT_ref) < TOL and rel(dA, dA_ref) < 5 * TOL
    if ini:
        assert rel(dh0, dh0_ref) < 5 * TOL
    assert all(np.isfinite(s).all() for s in sv) and np.all(dA[gm == 0] == 0)


def test_gru_row_order_is_a_pure_permutation(lib):
    """Length-sorted tiles (row_order) must give bit-identical results to the natural order."""
    B, W, G = 200, 30, 200
    XW, gm, Wh, h0, dhT = make(B, W, G, seed=3)
    order = np.argsort(-gm.sum(1), kind='stable').astype(np.int32)
    a = run(lib, 'cluster', XW, gm, Wh, h0, dhT, 'hard_sigmoid')
    b = run(lib, 'cluster', XW, gm, Wh, h0, dhT, 'hard_sigmoid', order=order)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize('B,W,G,ini,act,pad', [
    (1, 3, 8, True, 'hard_sigmoid', 'left'),
    (7, 5, 8, False, 'hard_sigmoid', 'holes'),
    (64, 50, 200, True, 'hard_sigmoid', 'left'),
    (50, 12, 104, False, 'sigmoid', 'holes'),
    (300, 20, 64, True, 'hard_sigmoid', 'left'),
    (1024, 50, 200, True, 'hard_sigmoid', 'left'),
    (130, 9, 224, True, 'sigmoid', 'left'),               # largest width whose weights fit tensor memory
    (40, 200, 64, True, 'hard_sigmoid', 'left'),          # C5 window
])
def test_gru_tensor_core_recurrence_vs_oracle(lib, B, W, G, ini, act, pad):
    """tcgen05 recurrence (3-term fp16 split of Wh and of the state, fp32 accumulate): fp32-like accuracy."""
    XW, gm, Wh, h0, dhT = make(B, W, G, seed=B + W + G, ini=ini, pad=pad)
    hT_ref, dA_ref, dh0_ref = oracle(XW, Wh, h0, dhT, act)
    hT, dA, dh0, sv = run(lib, 'tc', XW, gm, Wh, h0, dhT, act)
    assert rel(hT, hT_ref) < TOL
    assert rel(dA, dA_ref) < 5 * TOL
    if ini:
        assert rel(dh0, dh0_ref) < 5 * TOL
    assert all(np.isfinite(s).all() for s in sv)
    ref = run(lib, 'cluster', XW, gm, Wh, h0, dhT, act)
    for x, y in zip(sv, ref[3]):
        assert rel(x, y) < TOL


def test_gru_tensor_core_row_order(lib):
    B, W, G = 200, 30, 200
    XW, gm, Wh, h0, dhT = make(B, W, G, seed=3)
    order = np.argsort(-gm.sum(1), kind='stable').astype(np.int32)
    a = run(lib, 'tc', XW, gm, Wh, h0, dhT, 'hard_sigmoid')
    b = run(lib, 'tc', XW, gm, Wh, h0, dhT, 'hard_sigmoid', order=order)
    assert rel(b[0], a[0]) < TOL


@pytest.mark.parametrize('impl', ['tcb', 'tctc'])
@pytest.mark.parametrize('B,W,G,ini,act,pad,gscale', [
    (1, 3, 8, True, 'hard_sigmoid', 'left', 1.0),
    (7, 5, 8, False, 'hard_sigmoid', 'holes', 1e-6),       # tiny gradients: the per-tile scale keeps them in fp16 range
    (64, 50, 200, True, 'hard_sigmoid', 'left', 1.0 / 1024),
    (50, 12, 104, False, 'sigmoid', 'holes', 1e4),
    (300, 20, 64, True, 'hard_sigmoid', 'left', 1.0),
    (1024, 50, 200, True, 'hard_sigmoid', 'left', 1.0 / 1024),
    (130, 9, 224, True, 'sigmoid', 'left', 1.0),
    (40, 200, 64, True, 'hard_sigmoid', 'holes', 1.0),
])
def test_gru_tensor_core_bptt_vs_oracle(lib, impl, B, W, G, ini, act, pad, gscale):
    """tcgen05 BPTT: weights enter as fp16 (2^-12 relative rounding, like the other tensor-core-mode backward GEMMs), the
    exchanged gradients as fp16 hi+lo under a power-of-two scale that follows the cluster-wide max |d h| step by step
    (the W=200 case decays d h0 to 1e-22, far outside a fixed fp16 window).  Tolerance 2e-3 of the largest gradient at
    W <= 50; the weight rounding accumulates like sqrt(W) beyond."""
    XW, gm, Wh, h0, dhT = make(B, W, G, seed=B + W + G, ini=ini, pad=pad)
    dhT = dhT * gscale
    hT_ref, dA_ref, dh0_ref = oracle(XW, Wh, h0, dhT, act)
    hT, dA, dh0, sv = run(lib, impl, XW, gm, Wh, h0, dhT, act)
    tol = 2e-3 * max(1.0, (W / 50.0) ** 0.5)
    assert np.isfinite(dA).all()
    assert rel(dA, dA_ref) < tol
    if ini:
        assert rel(dh0, dh0_ref) < tol
    assert np.all(dA[gm == 0] == 0)
