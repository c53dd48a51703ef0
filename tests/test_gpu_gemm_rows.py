"""Row-listed tensor-core GEMMs (lstur_gemm_tc_tn_rows / lstur_gemm_tc_mrows): the products of the step whose reduction or
output rows are half history padding run over an ascending list of live rows with a device-side count.  Checked here
against torch on the same lists: empty list, full list, ragged sizes, bias / relu / transposed B, and that rows outside
the list are left untouched."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
P_ = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _rows(n, frac, seed):
    g = np.random.default_rng(seed)
    live = np.nonzero(g.random(n) < frac)[0].astype(np.int32)
    idx = torch.zeros(n, dtype=torch.int32, device='cuda')
    idx[:len(live)] = torch.as_tensor(live).cuda()
    cnt = torch.tensor([len(live)], dtype=torch.int32, device='cuda')
    return live, idx, cnt


@pytest.mark.parametrize('M,N,K,frac', [(200, 600, 51200, 0.5), (200, 400, 4096, 0.0), (72, 136, 3000, 1.0), (400, 200, 7001 * 4, 0.37)])
def test_tn_rows_matches_dense_sum_over_listed_rows(lib, M, N, K, frac):
    torch.manual_seed(M + N)
    A = torch.randn(K, M, device='cuda')
    B = torch.randn(K, N, device='cuda')
    live, idx, cnt = _rows(K, frac, K)
    C = torch.full((M, N), 7.0, device='cuda')
    nb = lib.lstur_gemm_tc_workspace_bytes(M, N, K)
    ws = torch.empty(max(nb, 4), dtype=torch.uint8, device='cuda')
    assert lib.lstur_gemm_tc_tn_rows(M, N, K, P_(A), M, P_(B), N, P_(C), N, P_(idx), P_(cnt), P_(ws), nb, _st()) == 0, lib.lstur_last_error()
    torch.cuda.synchronize()
    sel = torch.as_tensor(live).long().cuda()
    ref = (A[sel].half().double().t() @ B[sel].half().double())          # operands are rounded to fp16, products exact
    scale = float(ref.abs().max()) if len(live) else 1.0
    assert float((C.double() - ref).abs().max()) <= 2e-3 * max(scale, 1.0)
    if len(live) == 0:
        assert not C.any()


@pytest.mark.parametrize('transB,flags,use_bias', [(0, 4, True), (0, 0, False), (1, 0, False), (0, 1 | 4, True)])
@pytest.mark.parametrize('M,N,K,frac', [(51200, 600, 200, 0.5), (1000, 200, 400, 0.0), (777, 72, 88, 1.0), (4099, 400, 200, 0.3)])
def test_mrows_writes_only_the_listed_rows(lib, M, N, K, frac, transB, flags, use_bias):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda')
    B = torch.randn((N, K) if transB else (K, N), device='cuda') * 0.2
    bias = torch.randn(N, device='cuda') if use_bias else None
    live, idx, cnt = _rows(M, frac, M + 1)
    C = torch.full((M, N), -3.0, device='cuda')
    nb = lib.lstur_gemm_tc_workspace_bytes(M, N, K)
    ws = torch.empty(max(nb, 4), dtype=torch.uint8, device='cuda')
    rc = lib.lstur_gemm_tc_mrows(transB, M, N, K, P_(A), K, P_(B), B.shape[1], P_(C), N, P_(bias), flags, P_(idx), P_(cnt), P_(ws), nb, _st())
    assert rc == 0, lib.lstur_last_error()
    torch.cuda.synchronize()
    sel = torch.as_tensor(live).long().cuda()
    Bm = B.t() if transB else B
    if flags & 4:        # 3-term split: ~fp32 accuracy
        ref = A[sel].double() @ Bm.double()
        tol = 2e-5
    else:
        ref = A[sel].half().double() @ Bm.half().double()
        tol = 2e-3
    if use_bias:
        ref = ref + bias.double()
    if flags & 1:
        ref = torch.relu(ref)
    got = C[sel].double()
    if len(live):
        assert float((got - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    dead = torch.ones(M, dtype=torch.bool, device='cuda')
    dead[sel] = False
    assert bool((C[dead] == -3.0).all())                               # rows outside the list are untouched
