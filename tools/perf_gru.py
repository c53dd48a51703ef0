"""Micro-benchmark of the GRU recurrence kernels at the C3 shape (run on the GPU box)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
from mnexp_b200 import _lib
lib = _lib.load()
P_ = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
B, W, G = 1024, 50, 200
g = np.random.default_rng(0)
lens = np.clip(g.geometric(1 / 30.0, B), 1, W)
gm = np.zeros((B, W), np.float32)
for b in range(B):
    gm[b, W - lens[b]:] = 1
print('active fraction %.3f' % gm.mean())
XW = torch.randn(B, W, 3 * G, device='cuda') * torch.tensor(gm).cuda()[:, :, None]
gmd = torch.tensor(gm).cuda()
Wh = torch.randn(G, 3 * G, device='cuda') / G ** 0.5
WhT = Wh.t().contiguous()
h0 = torch.rand(B, G, device='cuda') - 0.5
hT = torch.empty(B, G, device='cuda')
sv = [torch.empty(B, W, G, device='cuda') for _ in range(5)]
dA = torch.empty(B, W, 3 * G, device='cuda'); dh0 = torch.empty(B, G, device='cuda'); dhT = torch.randn(B, G, device='cuda')
order = torch.tensor(np.argsort(-gm.sum(1), kind='stable').astype(np.int32)).cuda()
full = torch.ones_like(gmd)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def chk(rc):
    assert rc == 0, lib.lstur_last_error()
for name, m, od in (('all steps active', full, None), ('padded, natural order', gmd, None), ('padded, length-sorted', gmd, order)):
    f_cl = lambda: chk(lib.lstur_gru_fwd_cluster(B, W, G, P_(XW), P_(m), P_(h0), G, P_(Wh), 0, P_(hT), G, *[P_(s) for s in sv], P_(od), st()))
    b_cl = lambda: chk(lib.lstur_gru_bwd_cluster(B, W, G, P_(m), *[P_(s) for s in sv[:4]], P_(WhT), 0, P_(dhT), G, P_(dA), P_(dh0), G, P_(od), st()))
    f_st = lambda: chk(lib.lstur_gru_fwd_streaming(B, W, G, P_(XW), P_(m), P_(h0), G, P_(Wh), 0, P_(hT), G, *[P_(s) for s in sv], st()))
    b_st = lambda: chk(lib.lstur_gru_bwd_streaming(B, W, G, P_(m), *[P_(s) for s in sv[:4]], P_(WhT), 0, P_(dhT), G, P_(dA), P_(dh0), G, st()))
    f_tc = lambda: chk(lib.lstur_gru_fwd_tc(B, W, G, P_(XW), P_(m), P_(h0), G, P_(Wh), 0, P_(hT), G, *[P_(s) for s in sv], P_(od), st()))
    b_tc = lambda: chk(lib.lstur_gru_bwd_tc(B, W, G, P_(m), *[P_(s) for s in sv[:4]], P_(Wh), 0, P_(dhT), G, P_(dA), P_(dh0), G, P_(od), None, st()))
    print('%-24s tensor-core fwd %7.1f us bwd %7.1f us' % (name, timeit(f_tc), timeit(b_tc)))
    print('%-24s cluster fwd %7.1f us bwd %7.1f us | streaming fwd %7.1f us bwd %7.1f us' % (name, timeit(f_cl), timeit(b_cl), timeit(f_st), timeit(b_st)))
