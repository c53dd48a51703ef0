"""Data parallel == single GPU (mnexp_b200/dist.py): N replicas on the shards of a batch, after the exchange step,
apply exactly the update one engine applies on the whole batch (to fp32 reassociation), and stay bit-identical to each
other.  (1) two engines emulate two ranks on ONE GPU with the collectives done by hand; (2) a real 2-rank NCCL run
(skipped with fewer than 2 GPUs).  Dropout is 0: the step seed depends on (step, rank)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from mnexp_b200 import synth
from tolerances import assert_adam_weights_close, rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('precision,trainable', [('fp32', False), ('fp16_tc', False), ('fp32', True), ('fp16_tc', True)])
def test_two_emulated_ranks_equal_one_engine(lib, precision, trainable):
    from mnexp_b200.dist import UserRowReducer, shard_batch
    from mnexp_b200.engine import LsturEngine
    world, steps = 2, 3
    sh = synth.Shape('dp', 40, 300, 2000, L=30, W=50, K=4, B=16, E=300, F=400, U=200)
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=5)
    batches, _ = synth.make_batches(sh, steps, seed=6)
    kw = dict(arch='igru', doc_tokens=tok, dropout=0.0, lr=1e-3, precision=precision, trainable_word_emb=trainable)
    ranks = [LsturEngine(P, sh.B // world, sh.W, 1 + sh.K, sh.L, **kw) for _ in range(world)]
    full = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, **kw)
    red = [UserRowReducer(e, sh.B) for e in ranks]
    first = True
    for b in batches:
        dbs = [e.to_device_batch(shard_batch(b, r, world)) for r, e in enumerate(ranks)]
        for e, db in zip(ranks, dbs):
            e.step_seed += 1
            e.forward(db, training=True, seed=e.step_seed)
            e.backward(db, grad_scale=1.0 / sh.B)
        # the exchange of mnexp_b200.dist.exchange, by hand: all-reduce(sum) of the dense arenas (and of the dense
        # word-table gradient), rank-major all-gather of (user id, d user row)
        dense = ranks[0].dense_grad + ranks[1].dense_grad
        ids = torch.cat([db['user'] for db in dbs])
        rows = torch.cat([e.view('d_u0').reshape(e.B, e.Ue) for e in ranks]).contiguous()
        wsum = (ranks[0].word_grad + ranks[1].word_grad) if trainable else None
        if first:
            # the exchanged gradients of the first step == the gradients of one engine on the whole batch (before any
            # optimizer step can amplify rounding differences): fp32 to reassociation, tensor-core mode to the 16-bit
            # rounding of the saved activations / gradient images (tile-dependent scales in the recurrence)
            first = False
            full.step_seed += 1
            dbf = full.to_device_batch(b)
            full.forward(dbf, training=True, seed=full.step_seed)
            full.backward(dbf)
            gtol = 2e-6 if precision == 'fp32' else 2e-4
            assert rel(dense.cpu().numpy(), full.dense_grad.cpu().numpy()) < gtol
            if trainable:
                assert rel(wsum.cpu().numpy(), full.word_grad.cpu().numpy()) < gtol
            full.apply_adam()
        else:
            full.train_step(full.to_device_batch(b))
        for e, r in zip(ranks, red):
            e.dense_grad.copy_(dense)
            if trainable:
                e.word_grad.copy_(wsum)
            e.apply_adam(user_rows=r.reduce(e, ids, rows))
    torch.cuda.synchronize()
    w0, w1, wf = ranks[0].get_weights_dict(), ranks[1].get_weights_dict(), full.get_weights_dict()
    # fp32: reassociation only, every element.  Tensor-core mode: the 16-bit operand copies round-trip through the exchange
    # bit-exactly while the table is frozen; with the table trainable a 1e-7 reassociation difference in a word row can
    # flip its 16-bit rounding in the next step's operand copy.  Adam then normalises every gradient to ~lr: the few
    # elements whose gradient is rounding noise take +-lr steps of noise-decided sign (tolerances.py), so the bound is
    # "all but 1e-3 of the elements within tol, none beyond 2*lr*steps" (measured: 1 element of conv_w at 1.1e-3)
    tol = 1e-6 if precision == 'fp32' else (3e-4 if trainable else 1e-4)
    for k in wf:
        assert np.array_equal(w0[k], w1[k]), k                 # replicas bit-identical
        assert_adam_weights_close(w0[k], wf[k], tol, 1e-3, steps, name=k, exact=(precision == 'fp32'), tail=1e-3)
    if trainable:
        assert np.abs(wf['word_emb'] - P['word_emb']).max() > 1e-4      # the table actually moved


@pytest.mark.parametrize('precision,trainable', [('fp32', 0), ('fp16_tc', 0), ('fp16_tc', 1)])
def test_two_nccl_ranks_equal_one_engine(lib, precision, trainable):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29731', os.path.join(ROOT, 'tests', 'dp_worker.py'), precision, str(trainable)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize('precision', ['fp32', 'fp16_tc'])
def test_drop_in_surface_data_parallel(lib, precision):
    """torchrun + the task handler / Keras-protocol model (keras_like routes train steps through dist.DataParallel when a
    process group exists): two ranks on the halves of the reference-run batch reproduce the update the reference's own graph
    made on the whole batch (tests/dp_surface_worker.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29741', os.path.join(ROOT, 'tests', 'dp_surface_worker.py'), precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
