#!/usr/bin/env python
"""Generates tests/golden/lstur_golden.npz from the float64 oracle on seeded synthetic inputs.

The reference ships no golden vectors for this path and its Keras/TF-1.x runtime cannot be imported here
(SURVEY.md §8c), so these fixtures pin the ORACLE (and through it every CUDA kernel) against regressions; they
are not outputs of the reference itself (those are tests/golden/ref_golden.npz, make_ref_golden.py).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mnexp_b200 import synth  # noqa: E402
from oracle import lstur_numpy as on  # noqa: E402
from oracle import lstur_torch as ot  # noqa: E402


def main():
    out = {}
    sh = synth.SHAPES['tiny']
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    out['doc_tokens'] = tok
    for arch in ('igru', 'gru', 'hgru', 'nigru'):
        P = synth.make_weights(sh, arch=arch, bias_noise=0.05, seed=4242)
        (b,), _ = synth.make_batches(sh, 1, seed=99)
        ct, cd = tok[b['hist_doc']], tok[b['cand_doc']]
        r = on.lstur_forward(P, b['user'], ct, cd, arch=arch, aux=True)
        ora = ot.LsturOracle(P, arch=arch)
        loss, grads = ora.loss_and_grads(b['user'], ct, cd)
        for k in ('probs', 'logits', 'sigmoid', 'user_vec', 'cand_vec', 'hist_vec'):
            out['%s/%s' % (arch, k)] = r[k]
        out['%s/loss' % arch] = np.float64(loss)
        for k, g in grads.items():
            out['%s/grad/%s' % (arch, k)] = g.numpy()
        for k in ('user', 'hist_doc', 'cand_doc'):
            out['%s/batch/%s' % (arch, k)] = b[k]
        # 2 Keras-Adam steps (dense) from these weights
        l1 = ora.train_step(b['user'], ct, cd, training=False)
        l2 = ora.train_step(b['user'], ct, cd, training=False)
        out['%s/adam_losses' % arch] = np.array([l1, l2])
        out['%s/adam_conv_w' % arch] = ora.P['conv_w'].detach().numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lstur_golden.npz'), **out)
    print('wrote %d arrays' % len(out))


if __name__ == '__main__':
    main()
