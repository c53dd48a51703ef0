"""Ablation matrix of the tensor-core news-encoder forward (run on the GPU box): LSTUR_FWD_DBG bit combinations in one
process (the flag is read at every launch); the stage depth comes from LSTUR_FWD_STAGES (read once)."""
import ctypes, os, sys
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
flop = N * L * 2 * 3 * E * F
drop = float(os.environ.get("DROP", "0.2"))
call = lambda: lib.lstur_news_conv_tc_fwd(N, L, E, F, V, P_(tok), P_(emb), P_(wimg), P_(cb), P_(aw), P_(ab), P_(c_out), P_(pooled), P_(a), P_(w), ctypes.c_float(drop), 1, 1, 0, st())
for dbg in [int(x) for x in (sys.argv[1:] or ['0', '1', '3', '19', '35', '51'])]:
    os.environ['LSTUR_FWD_DBG'] = str(dbg)
    for i in range(7):
        if i == 2:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        assert call() == 0, lib.lstur_last_error()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    trace = torch.zeros(8 * 16, dtype=torch.int64, device='cuda')
    lib.lstur_tc_set_trace(P_(trace)); call(); torch.cuda.synchronize(); lib.lstur_tc_set_trace(None)
    tr = trace.cpu().numpy()
    print('stages=%s dbg=%2d  %.3f ms  %4.0f TFLOP/s   issuer waits: accumulator %5.1f%%  full stages %5.1f%%  (%d cycles)' % (
        os.environ.get('LSTUR_FWD_STAGES', 'def'), dbg, ms, flop / ms / 1e9, 100.0 * tr[0] / tr[2], 100.0 * tr[1] / tr[2], tr[2]))
