"""Micro-benchmark of the tensor-core news-encoder kernels (run on the GPU box)."""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from mnexp_b200 import _lib, synth
lib = _lib.load()
P_ = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
N, L, E, F, V = 56320, 30, 300, 400, 100000
g = np.random.default_rng(0)
tokd, _, _ = synth.make_docs(130000, L, V)
tok = torch.as_tensor(tokd[g.integers(0, 130001, N)]).cuda()
Ep = lib.lstur_tc_padded_e(E)
emb = (torch.randn(V, Ep, device='cuda') * 0.1).half()
wimg = (torch.randn(lib.lstur_tc_wimg_elems(E, F), device='cuda') * 0.05).half()
cb, aw, ab = torch.zeros(F, device='cuda'), torch.randn(F, device='cuda') * 0.1, torch.zeros(1, device='cuda')
c_out = torch.empty((N, L, F), dtype=torch.float16, device='cuda')
pooled = torch.empty((N, F), device='cuda'); a = torch.empty((N, L), device='cuda'); w = torch.empty((N, L), device='cuda')
def run(drop, ctas, reps=5):
    for i in range(reps + 2):
        if i == 2:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        rc = lib.lstur_news_conv_tc_fwd(N, L, E, F, V, P_(tok), P_(emb), P_(wimg), P_(cb), P_(aw), P_(ab), P_(c_out), P_(pooled), P_(a), P_(w), ctypes.c_float(drop), 1, 1, ctas, st())
        assert rc == 0, lib.lstur_last_error()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
flop = N * L * 2 * 3 * E * F
for drop in (0.0, 0.2):
    for ctas in (0, 111, 74, 37):
        ms = run(drop, ctas)
        print('fwd dropout=%.1f ctas=%3d  %.3f ms  %.0f TFLOP/s' % (drop, ctas or 148, ms, flop / ms / 1e9))

trace = torch.zeros(8 * 16, dtype=torch.int64, device='cuda')

# ---- wgrad
nb = lib.lstur_tc_dpre_img_bytes(N, L, F)
img = (torch.randn(nb // 2, device='cuda') * 0.01).half()
pb = lib.lstur_tc_wgrad_partial_bytes(N, E, F)
ws = torch.empty(pb, dtype=torch.uint8, device='cuda')
dW = torch.empty((3, E, F), device='cuda')
def runw(drop, reps=5):
    for i in range(reps + 2):
        if i == 2:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        rc = lib.lstur_conv_wgrad_tc(N, L, E, F, V, P_(tok), P_(emb), P_(img), P_(dW), ctypes.c_float(drop), 1, 1, P_(ws), pb, st())
        assert rc == 0, lib.lstur_last_error()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for drop in (0.0, 0.2):
    ms = runw(drop)
    print('wgrad dropout=%.1f %.3f ms %.0f TFLOP/s' % (drop, ms, flop / ms / 1e9))
    trace.zero_(); lib.lstur_tc_set_trace(P_(trace)); runw(drop, 1); lib.lstur_tc_set_trace(None)
    tr = trace.cpu().numpy()
    print('  wgrad trace: mma wait_full=%d of %d cycles over %d blocks; producer wait_empty=%d of %d; loader wait_empty=%d' % tuple(tr[:6][[0,1,2,3,4,5]]))
