"""Generates tests/golden/train_parity_c1.npz: the ORACLE arm of the training-parity check (north_star: "AUC within
0.002 after a fixed step count").

Arm = oracle.lstur_torch (fp32, dense Keras-Adam on every tensor incl. the user table = reference semantics,
task/paper.py:656) trained for K steps at BASELINE config C1 (LSTUR-ini, B=64) on the seeded learnable task of
mnexp_b200.synth.make_preference_task, then scored on held-out impressions (per-impression AUC, task/paper.py:504-515).
Two runs: dropout 0, and dropout 0.2 with the device's counter-based dropout stream replayed mask-for-mask
(mnexp_b200/rng.py::tc_dropout_multipliers; seed of step s = s, X stream 2s, C stream 2s+1 — LsturEngine.train_step).
The GPU test (tests/test_gpu_training_parity.py) trains the engine (fp16_tc, row-sparse and dense Adam) on the same
data from the same initial weights and compares AUCs.  Run here (CPU container): python tests/golden/make_train_parity.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mnexp_b200 import rng, synth          # noqa: E402
from oracle import lstur_torch as ot       # noqa: E402

K_STEPS, N_EVAL, LR, SEED = 200, 32, 1e-3, 4321
EP = 320                                    # lstur_tc_padded_e(300): the X-dropout stream is indexed over padded rows


def task():
    sh = synth.SHAPES['C1']
    tok, word_emb, train, evalb = synth.make_preference_task(sh, K_STEPS, N_EVAL, seed=SEED)
    P = synth.make_weights(sh, arch='igru', seed=99, word_emb=word_emb)
    return sh, tok, P, train, evalb


def quad_masks(seed, toks, E, F, p):
    """the engine's tensor-core dropout streams of step `seed` (indexed by the compacted title index, mnexp_b200/rng.py)"""
    mx, mc = rng.tc_dropout_multipliers(seed, toks, E, EP, F, p, dtype=np.float32)
    return torch.from_numpy(mx), torch.from_numpy(mc)


def train_step(ora, sh, tok, b, masks):
    P = ora.P
    user = torch.as_tensor(b['user']).long()
    ht = torch.as_tensor(tok[b['hist_doc']]).long()
    ct = torch.as_tensor(tok[b['cand_doc']]).long()
    B, W, L = ht.shape
    C = ct.shape[1]
    toks = torch.cat([ht.reshape(B * W, L), ct.reshape(B * C, L)])
    dx, dc = masks if masks is not None else (None, None)
    d = ot.news_encoder(toks, P, drop_x=dx, drop_c=dc)
    H = d[:B * W].reshape(B, W, -1) * (ht != 0).any(-1).to(d.dtype).unsqueeze(-1)
    u = ot.user_encoder('igru', user, H, P)
    probs = torch.softmax(ot.score(u, d[B * W:].reshape(B, C, -1)), -1)
    y = torch.zeros_like(probs)
    y[:, 0] = 1.0
    loss = ot.categorical_crossentropy(y, probs)
    gs = torch.autograd.grad(loss, [P[k] for k in ora.trainable], allow_unused=True)
    ora.opt.step(dict(zip(ora.trainable, gs)))
    return float(loss.detach())


def evaluate(ora, tok, evalb):
    out = []
    with torch.no_grad():
        for b in evalb:
            r = ora.forward(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], aux=True)
            out.append(r['logits'].numpy().astype(np.float32))
    return np.concatenate(out)


def run(p):
    sh, tok, P, train, evalb = task()
    torch.manual_seed(0)
    ora = ot.LsturOracle(P, arch='igru', dtype=torch.float32, lr=LR)
    losses = []
    t0 = time.time()
    for s, b in enumerate(train, 1):
        toks = np.concatenate([tok[b['hist_doc']].reshape(-1, sh.L), tok[b['cand_doc']].reshape(-1, sh.L)])
        masks = quad_masks(s, toks, sh.E, sh.F, p) if p > 0 else None
        losses.append(train_step(ora, sh, tok, b, masks))
        if s % 20 == 0:
            print('p=%.1f step %d loss %.4f (%.0f s)' % (p, s, losses[-1], time.time() - t0), flush=True)
    logits = evaluate(ora, tok, evalb)
    return np.asarray(losses, dtype=np.float32), logits, evaluate_init(P, tok, evalb)


def evaluate_init(P, tok, evalb):
    ora = ot.LsturOracle(P, arch='igru', dtype=torch.float32)
    return evaluate(ora, tok, evalb)


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    out = {}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'train_parity_c1.npz')
    only = sys.argv[1:]                       # e.g. `p2`: regenerate one run, keep the others from the existing file
    if only and os.path.exists(path):
        out = dict(np.load(path))
    for name, p in (('p0', 0.0), ('p2', 0.2)):
        if only and name not in only:
            continue
        losses, logits, logits0 = run(p)
        out['loss_' + name], out['logits_' + name] = losses, logits
        out['auc_' + name] = np.float64(synth.impression_auc(logits))
        out['auc_init'] = np.float64(synth.impression_auc(logits0))
        print(name, 'AUC', out['auc_' + name], 'init', out['auc_init'], flush=True)
    out['k_steps'], out['n_eval'], out['lr'], out['seed'] = K_STEPS, N_EVAL, LR, SEED
    np.savez_compressed(path, **out)
