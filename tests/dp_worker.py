"""Worker of tests/test_gpu_dp.py::test_two_nccl_ranks_equal_one_engine — launched with torch.distributed.run, one
process per GPU.  Every rank trains STEPS data-parallel steps on its shard of a global batch (dropout 0), then rank 0
trains one engine on the concatenated batch and compares; exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mnexp_b200 import synth                                   # noqa: E402
from mnexp_b200.dist import DataParallel, shard_batch          # noqa: E402
from mnexp_b200.engine import LsturEngine                      # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from tolerances import assert_adam_weights_close               # noqa: E402

STEPS = 3


def main():
    precision, trainable = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    from mnexp_b200.dist import init_process_group
    init_process_group(int(os.environ['LOCAL_RANK']))
    sh = synth.Shape('dp', 40, 300, 2000, L=30, W=50, K=4, B=8 * world, E=300, F=400, U=200)
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch='igru', bias_noise=0.05, seed=5)
    batches, _ = synth.make_batches(sh, STEPS, seed=6)
    kw = dict(arch='igru', doc_tokens=tok, dropout=0.0, lr=1e-3, precision=precision, trainable_word_emb=bool(trainable))
    eng = LsturEngine(P, sh.B // world, sh.W, 1 + sh.K, sh.L, **kw)
    dp = DataParallel(eng)
    for b in batches:
        dp.train_step(eng.to_device_batch(shard_batch(b, rank, world)))
    torch.cuda.synchronize()
    w = eng.get_weights_dict()
    # replicas must be bit-identical: compare a checksum of every tensor across ranks
    names = sorted(w)
    sums = torch.tensor([float(np.asarray(w[k], dtype=np.float64).sum()) for k in names], dtype=torch.float64, device='cuda')
    allsums = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(allsums, sums)
    ok = all(torch.equal(allsums[0], s) for s in allsums)
    if rank == 0:
        full = LsturEngine(P, sh.B, sh.W, 1 + sh.K, sh.L, **kw)
        for b in batches:
            full.train_step(full.to_device_batch(b))
        torch.cuda.synchronize()
        wf = full.get_weights_dict()
        tol = 1e-6 if precision == 'fp32' else (3e-4 if trainable else 1e-4)
        for k in names:
            try:            # tolerances.py: all but 1e-3 of the elements within tol, none beyond 2*lr*steps (fp32: every element)
                err, frac = assert_adam_weights_close(w[k], wf[k], tol, 1e-3, STEPS, name=k, exact=(precision == 'fp32'))
            except AssertionError as e:
                print('MISMATCH', e)
                ok = False
                continue
            print('%-10s max |dp - single| = %.3e (%.1e of the elements above %.0e)' % (k, err, frac, tol))
        print('replicas identical and equal to the single-GPU run:', ok)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
