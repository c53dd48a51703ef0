"""Worker of tests/test_gpu_dp.py::test_drop_in_surface_data_parallel — launched with torch.distributed.run, one process per
GPU.  Every rank builds the mirror's Seq2VecPaperSoftmaxId handler with config.batch_size = 3 and trains three
train_on_batch steps on ITS half of the six-row batch of the reference-run case `sid-igru-dot`; the exchanged update must
be the one the reference's own graph made on the whole batch (tests/golden/ref_golden.npz: losses and weights after three
Adam steps), and the replicas must stay bit-identical."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from mnexp_b200 import settings, synth, task                   # noqa: E402
from mnexp_b200.dist import init_process_group                 # noqa: E402


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    init_process_group(int(os.environ['LOCAL_RANK']))
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_golden.npz'))
    name, sh = 'sid-igru-dot', synth.SHAPES['tiny']
    assert sh.B % world == 0
    per = sh.B // world
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sh)
    cfg = settings.Config(dict(task='Seq2VecPaperSoftmaxId', arch='igru', score_model='dot', input_training_data_path=d,
                               title_shape=sh.L, window_size=sh.W, negative_samples=sh.K, batch_size=per,
                               textual_embedding_dim=sh.E, title_filter_shape=(sh.F, sh.k), user_embedding_dim=sh.U, debug=True,
                               dropout=0.0, precision=precision, learning_rate=0.001, sparse_user_adam=False))
    h = task.get(cfg)
    model = h.build_model(0)
    P = {k[len(name) + 3:]: gold[k] for k in gold.files if k.startswith(name + '/P/')}
    names = [k for k in model.WEIGHT_ORDER if k in model._current()]
    model.set_weights([np.asarray(P[k], dtype=np.float32) for k in names])
    x = [gold['%s/x%d' % (name, i)][rank * per:(rank + 1) * per] for i in range(int(gold[name + '/n_inputs']))]
    y = gold[name + '/y'][rank * per:(rank + 1) * per]
    losses = [float(model.train_on_batch(x, y)[0]) for _ in range(3)]
    torch.cuda.synchronize()
    w = dict(zip(names, model.get_weights()))
    ok = True
    ref_losses = gold[name + '/adam_losses']
    tol = 1e-4 if precision == 'fp32' else 2e-3
    if np.abs(np.array(losses) - ref_losses).max() >= tol:
        print('rank %d: losses %s vs the reference graph %s' % (rank, losses, ref_losses))
        ok = False
    if precision == 'fp32':
        dd = np.concatenate([np.abs(np.asarray(w[k], dtype=np.float64).reshape(-1) - gold['%s/adam/%s' % (name, k)].reshape(-1))
                             for k in names])
        if not (np.mean(dd > 2e-5) <= 2e-3 and dd.max() <= 6.1e-3):
            print('rank %d: weights differ from the reference graph: max %.3e, %.2e of the elements above 2e-5'
                  % (rank, dd.max(), np.mean(dd > 2e-5)))
            ok = False
    sums = torch.tensor([float(np.asarray(w[k], dtype=np.float64).sum()) for k in names], dtype=torch.float64, device='cuda')
    allsums = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(allsums, sums)
    if not all(torch.equal(allsums[0], s) for s in allsums):
        print('rank %d: replicas differ' % rank)
        ok = False
    flag = torch.tensor([1.0 if ok else 0.0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('drop-in surface, %d ranks, %s: losses %s (reference %s): %s' % (world, precision, losses, list(ref_losses),
                                                                               'ok' if flag.item() == 1.0 else 'MISMATCH'))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
