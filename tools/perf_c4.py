"""BASELINE config 4: encode all 130k news once, then score 1M users x 20 candidates through the GRU user encoder
(decomposed pipeline, lstur_encode_docs + lstur_forward_docvecs).  Run on the GPU box."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from mnexp_b200 import synth
from mnexp_b200.engine import LsturEngine
sh = synth.SHAPES['C3']
n_users_scored = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B, C = 2048, 20
tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
P = synth.make_weights(sh, arch='igru')
eng = LsturEngine(P, B, sh.W, C, sh.L, arch='igru', doc_tokens=tok, precision='fp16_tc', training=False)
def timed(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)
table, ms_docs = timed(eng.build_doc_table)
table, ms_docs = timed(eng.build_doc_table)          # second pass: warm
print('news pass: %d docs in %.1f ms  -> %.0f docs/s' % (table.shape[0], ms_docs, table.shape[0] / ms_docs * 1e3))
g = np.random.default_rng(0)
n_batches = 8
dbs = []
for i in range(n_batches):
    lens = np.clip(g.geometric(1 / 30.0, B), 1, sh.W)
    hist = np.zeros((B, sh.W), np.int32)
    for r in range(B):
        hist[r, sh.W - lens[r]:] = g.integers(1, sh.n_news + 1, lens[r])
    dbs.append(eng.to_device_batch(dict(user=g.integers(0, sh.n_users, B).astype(np.int32), hist_doc=hist,
                                        cand_doc=g.integers(1, sh.n_news + 1, (B, C)).astype(np.int32))))
steps = (n_users_scored + B - 1) // B
def users():
    for i in range(steps):
        eng.forward_docvecs(dbs[i % n_batches], table)
        s = eng.score_sigmoid()
    return s
users()
_, ms_users = timed(users)
print('user pass: %d users x %d candidates in %.1f ms -> %.0f users/s, %.0f pairs/s' % (steps * B, C, ms_users, steps * B / ms_users * 1e3, steps * B * C / ms_users * 1e3))
print('C4 total: %.1f ms' % (ms_docs + ms_users))
