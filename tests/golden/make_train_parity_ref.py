#!/usr/bin/env python
"""Generates tests/golden/train_parity_c1_ref.npz: the training-parity run of make_train_parity.py (C1, 200 steps, dropout 0,
seeded learnable task) repeated by the REFERENCE'S OWN graph — /root/reference's unmodified task/paper.py
(Seq2VecPaperSoftmaxId._build_model, its compiled categorical cross-entropy and keras.optimizers.Adam) imported over
oracle/keras_shim in float32, trained with model.train_on_batch on the same batches from the same weights.

tests/test_ref_pinned.py::test_training_parity_fixture_matches_reference_graph then requires the oracle-trained fixture
(train_parity_c1.npz, what the CUDA arms are compared with on the GPU) to agree with it: per-step losses and the held-out
AUC.  Run in the build container (needs /root/reference):  python tests/golden/make_train_parity_ref.py
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_golden as mk          # noqa: E402
import make_train_parity as mtp       # noqa: E402


def main():
    keras, settings, task = mk.load_reference()
    keras.backend.set_floatx('float32')
    from mnexp_b200 import synth
    sh, tok, P, train, evalb = mtp.task()
    d = tempfile.mkdtemp(prefix='trainparity_')
    synth.write_docmeta_tsv(os.path.join(d, 'DocMeta.tsv'), tok)
    synth.write_clickdata_tsv(os.path.join(d, 'ClickData.tsv'), sh.n_users, sh.n_news, np.random.default_rng(1))
    np.save(os.path.join(d, 'Vocab.tsv.npy'), P['word_emb'])
    cfg = mk.reference_config(settings, d, sh, 'Seq2VecPaperSoftmaxId', 'igru', 'dot', learning_rate=mtp.LR)
    h = task.get(cfg)
    model = h.build_model(0)
    mk.assign(mk.named_variables(h), P)

    def feed(b):
        y = np.zeros((len(b['user']), 1 + sh.K), dtype=np.float32)
        y[:, 0] = 1.0
        return [b['user'], tok[b['hist_doc']]] + [tok[b['cand_doc'][:, j]] for j in range(1 + sh.K)], y
    losses, t0 = [], time.time()
    for s, b in enumerate(train, 1):
        x, y = feed(b)
        losses.append(model.train_on_batch(x, y)[0])
        if s % 20 == 0:
            print('step %d loss %.4f (%.0f s)' % (s, losses[-1], time.time() - t0), flush=True)
    probs = np.concatenate([model.predict(feed(b)[0], batch_size=sh.B) for b in evalb]).astype(np.float32)
    out = dict(ref_loss_p0=np.asarray(losses, dtype=np.float32), ref_probs_p0=probs,
               ref_auc_p0=np.float64(synth.impression_auc(probs)))
    np.savez_compressed(os.path.join(HERE, 'train_parity_c1_ref.npz'), **out)
    print('reference-graph AUC after %d steps: %.6f' % (len(losses), out['ref_auc_p0']))


if __name__ == '__main__':
    import torch
    torch.set_num_threads(os.cpu_count())
    main()
