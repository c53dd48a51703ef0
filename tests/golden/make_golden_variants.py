#!/usr/bin/env python
"""Generates tests/golden/lstur_golden_variants.npz: the remaining user encoders / scorers / the sigmoid family, from the
float64 torch oracle on seeded synthetic inputs (same status as lstur_golden.npz: a regression pin of the oracle; the vectors produced by the reference's own code are
tests/golden/ref_golden.npz).
Run from the repo root:  python tests/golden/make_golden_variants.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mnexp_b200 import synth  # noqa: E402
from oracle import lstur_torch as ot  # noqa: E402

CASES = (('ngru', 'dnn', 'softmax'), ('igru', 'ddot', 'softmax'), ('niavg', 'dnn', 'softmax'), ('iigru', 'dot', 'softmax'),
         ('igru', 'dnn', 'bce'), ('iicat', 'dnn', 'bce'))
GAIN = 1.5


def run_case(arch, score_model, head, sh, tok):
    P = synth.make_weights(sh, arch=arch, bias_noise=0.05, seed=5150, score_model=score_model)
    (b,), _ = synth.make_batches(sh, 1, seed=51)
    ora = ot.LsturOracle(P, arch=arch, score_model=score_model)
    cand = b['cand_doc'] if head == 'softmax' else b['cand_doc'][:, :1]
    u, c, d = ora._ints(b['user'], tok[b['hist_doc']], tok[cand])
    if head == 'softmax':
        out = ot.forward(ora.P, u, c, d, arch=arch, score_model=score_model)
        loss = ot.loss_fn(ora.P, u, c, d, arch=arch, score_model=score_model)
        y = None
    else:
        y = (np.random.default_rng(52).random((sh.B, 1)) < 0.4).astype(np.float32)
        out = ot.forward(ora.P, u, c, d, arch=arch, score_model=score_model, head='sigmoid')
        loss = ot.weighted_bce(torch.tensor(y, dtype=torch.float64), out, gain=GAIN, negative_samples=sh.K)
    grads = dict(zip(ora.trainable, torch.autograd.grad(loss, [ora.P[k] for k in ora.trainable], allow_unused=True)))
    return P, b, cand, y, out.detach().numpy(), float(loss.detach()), grads


def main():
    out = {}
    sh = synth.SHAPES['tiny']
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    for arch, score_model, head in CASES:
        key = '%s-%s-%s' % (arch, score_model, head)
        P, b, cand, y, probs, loss, grads = run_case(arch, score_model, head, sh, tok)
        out[key + '/probs'] = probs
        out[key + '/loss'] = np.float64(loss)
        if y is not None:
            out[key + '/label'] = y
        for k, g in grads.items():
            if g is not None and k != 'word_emb':
                out[key + '/grad/' + k] = g.numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lstur_golden_variants.npz'), **out)
    print('wrote %d arrays' % len(out))


if __name__ == '__main__':
    main()
