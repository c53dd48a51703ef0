"""Training parity (BASELINE.json north_star: "AUC within 0.002 after a fixed step count").

The oracle arm (oracle.lstur_torch: fp32, dense Keras-Adam = reference semantics) was trained offline for 200 steps at
config C1 on the seeded learnable task and its held-out logits are the committed fixture tests/golden/train_parity_c1.npz
(generator: tests/golden/make_train_parity.py).  Here the engine is trained on the same batches from the same initial
weights — tensor-core precision with the row-sparse user-table Adam (the throughput configuration bench.py runs), with
dense Adam (reference semantics), and in fp32 — and must rank the held-out impressions the same:
|mean per-impression AUC - oracle's| <= 0.002 (task/paper.py:504-515).  The dropout-0.2 run replays the device's
counter-based masks, which the fixture generator fed to the oracle mask-for-mask.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import make_train_parity as mtp          # noqa: E402
from mnexp_b200 import synth             # noqa: E402

pytestmark = pytest.mark.gpu
AUC_TOL = 0.002
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'train_parity_c1.npz')


@pytest.fixture(scope='module')
def fixture():
    return np.load(FIX)


@pytest.fixture(scope='module')
def data():
    return mtp.task()


_runs = {}


def run_engine(lib, data, precision, sparse, dropout):
    key = (precision, sparse, dropout)
    if key in _runs:
        return _runs[key]
    from mnexp_b200.engine import LsturEngine
    sh, tok, P, train, evalb = data
    eng = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, arch='igru', doc_tokens=tok, dropout=dropout, lr=mtp.LR,
                      precision=precision, sparse_user_adam=sparse)
    losses = [float(eng.train_step(eng.to_device_batch(b))[0]) for b in train]
    logits = []
    for b in evalb:
        eng.forward(eng.to_device_batch(b), training=False)
        logits.append(eng.view('logits').reshape(sh.B, -1).cpu().numpy().copy())
    _runs[key] = (np.asarray(losses), np.concatenate(logits))
    return _runs[key]


def test_fixture_is_a_learnable_task(fixture):
    """the check means something only if the oracle actually learnt to rank"""
    assert fixture['auc_p0'] > fixture['auc_init'] + 0.1
    assert fixture['auc_p2'] > fixture['auc_init'] + 0.1
    assert int(fixture['k_steps']) == mtp.K_STEPS == 200


@pytest.mark.parametrize('precision,sparse', [('fp16_tc', True), ('fp16_tc', False), ('fp32', False)])
def test_auc_after_200_steps_matches_oracle(lib, fixture, data, precision, sparse):
    losses, logits = run_engine(lib, data, precision, sparse, 0.0)
    auc = synth.impression_auc(logits)
    assert abs(auc - float(fixture['auc_p0'])) <= AUC_TOL, (auc, float(fixture['auc_p0']))
    # the first steps follow the oracle's loss to the precision's tolerance (later steps drift apart chaotically)
    tol = 1e-3 if precision == 'fp16_tc' else 1e-4
    assert np.abs(losses[:5] - fixture['loss_p0'][:5]).max() <= tol * 2.0
    # and the final smoothed loss agrees
    assert abs(losses[-20:].mean() - fixture['loss_p0'][-20:].mean()) <= 0.02


def test_auc_row_sparse_vs_dense_adam(lib, data):
    """the documented optimizer deviation (row-sparse Adam on the user table) does not change the ranking quality"""
    _, a = run_engine(lib, data, 'fp16_tc', True, 0.0)
    _, b = run_engine(lib, data, 'fp16_tc', False, 0.0)
    assert abs(synth.impression_auc(a) - synth.impression_auc(b)) <= AUC_TOL


def test_auc_with_dropout_masks_replayed(lib, fixture, data):
    losses, logits = run_engine(lib, data, 'fp16_tc', True, 0.2)
    auc = synth.impression_auc(logits)
    assert abs(auc - float(fixture['auc_p2'])) <= AUC_TOL, (auc, float(fixture['auc_p2']))
    assert np.abs(losses[:5] - fixture['loss_p2'][:5]).max() <= 2e-3
