"""Micro-benchmark of the attention-pooling backward (image output) at the C3 shape (run on the GPU box)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
from mnexp_b200 import _lib
lib = _lib.load()
P_ = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
N, L, F = 56320, 30, 400
C = (torch.randn(N, L, F, device='cuda').clamp_min(0) * 0.1).half()
a = torch.tanh(torch.randn(N, L, device='cuda')); w = torch.softmax(torch.randn(N, L, device='cuda'), -1)
dp = torch.randn(N, F, device='cuda'); ka = torch.randn(F, device='cuda') * 0.1
img = torch.empty(lib.lstur_tc_dpre_img_bytes(N, L, F), dtype=torch.uint8, device='cuda')
g = lib.lstur_attn_bwd_grid(N)
part = torch.empty(g * (2 * F + 1), device='cuda')
dka = torch.empty(F, device='cuda'); dcb = torch.empty(F, device='cuda'); dab = torch.empty(1, device='cuda')
f = lambda: lib.lstur_attn_pool_bwd_img(1, N, L, F, P_(C), P_(a), P_(w), P_(dp), F, P_(ka), P_(img), ctypes.c_float(0.2), ctypes.c_float(1.0), P_(dka), P_(dcb), P_(dab), 0, P_(part), part.numel() * 4, None, None, st())
for _ in range(3): assert f() == 0, lib.lstur_last_error()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
gb = (C.numel() * 2 + img.numel()) / 1e9
print('attn bwd img: grid %d  %.1f us  %.0f GB/s (C16 read + image write)' % (g, us, gb / us * 1e6))
