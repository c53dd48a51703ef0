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

    def feed(b):
        y = np.zeros((len(b['user']), 1 + sh.K), dtype=np.float32)
        y[:, 0] = 1.0
        return [b['user'], tok[b['hist_doc']]] + [tok[b['cand_doc'][:, j]] for j in range(1 + sh.K)], y
    from keras import _engine
    P0 = {k: np.array(v, copy=True) for k, v in P.items()}
    path = os.path.join(HERE, 'train_parity_c1_ref.npz')
    out = dict(np.load(path)) if os.path.exists(path) else {}
    only = sys.argv[1:]
    for name, p in (('p0', 0.0), ('p2', 0.2)):
        if only and name not in only:
            continue
        # dropout 0.2: the graph is rebuilt with Dropout(0.2) layers; their keep masks are the device's tensor-core stream of
        # step s (make_train_parity.quad_masks), fed through the shim's dropout hook in the order the graph evaluates them
        keras.backend.clear_session()
        cfg = mk.reference_config(settings, d, sh, 'Seq2VecPaperSoftmaxId', 'igru', 'dot', learning_rate=mtp.LR, dropout=p)
        h = task.get(cfg)
        model = h.build_model(0)
        mk.assign(mk.named_variables(h), P0)
        losses, t0 = [], time.time()
        for s, b in enumerate(train, 1):
            x, y = feed(b)
            if p > 0:
                toks = np.concatenate([tok[b['hist_doc']].reshape(-1, sh.L), tok[b['cand_doc']].reshape(-1, sh.L)])
                mx, mc = mtp.quad_masks(s, toks, sh.E, sh.F, p)
                _engine._STATE['dropout_hook'] = mk.MaskReplay(sh, keep={sh.E: mx.numpy() != 0, sh.F: mc.numpy() != 0})
            try:
                losses.append(model.train_on_batch(x, y)[0])
            finally:
                _engine._STATE['dropout_hook'] = None
            if s % 20 == 0:
                print('%s step %d loss %.4f (%.0f s)' % (name, s, losses[-1], time.time() - t0), flush=True)
        probs = np.concatenate([model.predict(feed(b)[0], batch_size=sh.B) for b in evalb]).astype(np.float32)
        out['ref_loss_' + name] = np.asarray(losses, dtype=np.float32)
        out['ref_probs_' + name] = probs
        out['ref_auc_' + name] = np.float64(synth.impression_auc(probs))
        print('%s: reference-graph AUC after %d steps: %.6f' % (name, len(losses), out['ref_auc_' + name]), flush=True)
    np.savez_compressed(path, **out)


if __name__ == '__main__':
    import torch
    torch.set_num_threads(os.cpu_count())
    main()
