"""wgrad-only micro-benchmark (run on the GPU box)."""
import ctypes, sys
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
nb = lib.lstur_tc_dpre_img_bytes(N, L, F)
img = (torch.randn(nb // 2, device='cuda') * 0.01).half()
pb = lib.lstur_tc_wgrad_partial_bytes(N, E, F)
ws = torch.empty(pb, dtype=torch.uint8, device='cuda')
dW = torch.empty((3, E, F), device='cuda')
trace = torch.zeros(8 * 16, dtype=torch.int64, device='cuda')
flop = N * L * 2 * 3 * E * F
xm = torch.randint(0, 256, (lib.lstur_tc_xmask_bytes(N, L, E),), dtype=torch.uint8, device='cuda')
def runm(drop, reps=5):
    for i in range(reps + 2):
        if i == 2:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        rc = lib.lstur_conv_wgrad_tc_m(N, L, E, F, V, P_(tok), P_(emb), P_(img), P_(dW), ctypes.c_float(drop), 1, 1, P_(ws), pb, P_(xm), ctypes.c_float(1.0), None, st())
        assert rc == 0, lib.lstur_last_error()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def runw(drop, reps=5):
    for i in range(reps + 2):
        if i == 2:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        rc = lib.lstur_conv_wgrad_tc(N, L, E, F, V, P_(tok), P_(emb), P_(img), P_(dW), ctypes.c_float(drop), 1, 1, P_(ws), pb, st())
        assert rc == 0, lib.lstur_last_error()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print('wgrad with keep bytes from the forward, dropout=0.2: %.3f ms' % runm(0.2))
trace.zero_(); lib.lstur_tc_set_trace(P_(trace)); runm(0.2, 1); lib.lstur_tc_set_trace(None)
tr = trace.cpu().numpy()
print('  producer (warp 4 of CTA 0) cycles: loads+mask %d, wait_empty %d, st.shared %d, proxy fence %d, syncwarp+arrive %d, total %d over %d stages'
      % (tr[6], tr[3], tr[7], tr[8], tr[9], tr[4], tr[2]))
for drop in (0.0, 0.2):
    ms = runw(drop)
    print('wgrad dropout=%.1f %.3f ms %.0f TFLOP/s' % (drop, ms, flop / ms / 1e9))
    trace.zero_(); lib.lstur_tc_set_trace(P_(trace)); runw(drop, 1); lib.lstur_tc_set_trace(None)
    tr = trace.cpu().numpy()
    print('  wgrad trace: mma wait_full=%d of %d cycles over %d blocks; producer wait_empty=%d of %d; loader wait_empty=%d' % tuple(tr[:6][[0,1,2,3,4,5]]))
