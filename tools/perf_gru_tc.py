"""One launch of the tensor-core GRU forward at the C3 shape (for the in-kernel phase counters, -DLSTUR_GRUTC_PROF)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
from mnexp_b200 import _lib
lib = _lib.load()
P_ = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
B, W, G = 1024, 50, 200
XW = torch.randn(B, W, 3 * G, device='cuda')
gm = torch.ones(B, W, device='cuda')
Wh = torch.randn(G, 3 * G, device='cuda') / G ** 0.5
h0 = torch.rand(B, G, device='cuda') - 0.5
hT = torch.empty(B, G, device='cuda')
sv = [torch.empty(B, W, G, device='cuda') for _ in range(5)]
save = len(sys.argv) < 2 or sys.argv[1] != 'nosave'
svp = [P_(s) if save else None for s in sv]
for _ in range(2):
    rc = lib.lstur_gru_fwd_tc(B, W, G, P_(XW), P_(gm), P_(h0), G, P_(Wh), 0, P_(hT), G, *svp, None, st())
    assert rc == 0, lib.lstur_last_error()
    torch.cuda.synchronize()
