"""Micro-benchmark of lstur_gemm_tc at the shapes of the LSTUR step (run on the GPU box)."""
import ctypes, sys
import torch
sys.path.insert(0, '.')
from mnexp_b200 import _lib
lib = _lib.load()
P_ = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def bench(ta, tb, M, N, K, flags, name, bias=False):
    A = torch.randn((K, M) if ta else (M, K), device='cuda')
    B = torch.randn((N, K) if tb else (K, N), device='cuda')
    C = torch.empty((M, N), device='cuda')
    bv = torch.randn(N, device='cuda') if bias else None
    nb = lib.lstur_gemm_tc_workspace_bytes(M, N, K)
    ws = torch.empty(max(nb, 4), dtype=torch.uint8, device='cuda')
    f = lambda: lib.lstur_gemm_tc(ta, tb, M, N, K, P_(A), A.shape[1], P_(B), B.shape[1], P_(C), N, P_(bv), flags, P_(ws), nb, st())
    for _ in range(3): assert f() == 0, lib.lstur_last_error()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    mb = (A.numel() + B.numel() + C.numel()) * 4 / 1e6
    print('%-28s ta=%d tb=%d %6dx%4dx%6d flags=%d  %7.1f us  %6.0f GB/s (operands+result once)' % (name, ta, tb, M, N, K, flags, us, mb / us * 1e3))
bench(0, 0, 56320, 200, 400, 4, 'dense fwd (precise)')
bench(0, 0, 56320, 200, 400, 0, 'dense fwd (single)')
bench(0, 0, 51200, 600, 200, 4, 'XW (precise)')
bench(0, 0, 51200, 600, 200, 4, 'XW (precise, bias)', bias=True)
bench(0, 0, 56320, 200, 400, 4, 'dense fwd (precise, bias)', bias=True)
bench(0, 0, 51200, 600, 200, 0, 'XW (single)')
bench(1, 0, 200, 600, 51200, 0, 'dWx')
bench(1, 0, 200, 400, 51200, 0, 'dWh zr')
bench(0, 1, 51200, 200, 600, 0, 'dH')
bench(0, 1, 56320, 400, 200, 0, 'd pooled')
bench(1, 0, 400, 200, 56320, 0, 'd dense_w')
