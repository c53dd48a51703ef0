"""Device-side batch assembly for the generator-fed LSTUR path.

The reference builds every training sample in Python (`train_gen` → `Window.get_title` → `np.stack`,
task/paper.py:396-405, task/seq2vec.py:17-53, 182-200), which caps its throughput far below the kernels.  Here the click
streams and the negatives of every impression live on the GPU as CSR tables, and one kernel (lstur_assemble_batch)
emits the compact doc-id batch `user (B) / hist_doc (B,W) / cand_doc (B,1+K)` that LsturEngine consumes:
  * history windows are bit-exact with `Window` (last W clicks before the sample, left-padded with doc 0);
  * the positive is the click itself, negatives are K draws with replacement from the impression's negatives
    (`Impression.negative_samples`), from the counter-based RNG instead of numpy's (distribution, not stream, parity);
  * sample order: a device permutation per epoch instead of the reference's 100·B shuffle pool.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def build_tables(data):
    """handler.data = [(train_impressions, valid_impressions)] per user -> host CSR tables of the training split."""
    stream_off, stream_docs, click_user, neg_off, neg_docs, sample_click = [0], [], [], [0], [], []
    for user, (ih, _) in enumerate(data):
        first = len(stream_docs)
        for imp in ih:
            for pos in imp.pos:
                c = len(stream_docs)
                if c > first:                           # `if ch.count:` — the first click of a user has no history
                    sample_click.append(c)
                stream_docs.append(int(pos))
                click_user.append(user)
                neg_docs.extend(int(x) for x in imp.neg)
                neg_off.append(len(neg_docs))
        stream_off.append(len(stream_docs))
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    return dict(stream_off=i32(stream_off), stream_docs=i32(stream_docs), click_user=i32(click_user), neg_off=i32(neg_off),
                neg_docs=i32(neg_docs if neg_docs else [0]), sample_click=i32(sample_click))


class DeviceBatcher:
    def __init__(self, data, B, W, K, device=None, seed=0):
        if not torch.cuda.is_available():
            raise _lib.LsturError('DeviceBatcher needs a CUDA device (no CPU fallback)')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.W, self.K, self.seed = B, W, K, seed
        self.host = build_tables(data)
        self.t = {k: torch.as_tensor(v).to(self.device) for k, v in self.host.items()}
        self.n_samples = int(self.host['sample_click'].shape[0])
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.perm, self.cursor, self.step = None, 0, 0

    def _next_indices(self):
        if self.perm is None or self.cursor + self.B > self.n_samples:
            self.perm = torch.randperm(self.n_samples, device=self.device, generator=self.gen, dtype=torch.int32)
            self.cursor = 0
        idx = self.perm[self.cursor:self.cursor + self.B].contiguous()
        self.cursor += self.B
        return idx

    def assemble(self, idx):
        """idx: (B) int32 device tensor of sample numbers -> device batch dict for LsturEngine."""
        B, W, K = int(idx.numel()), self.W, self.K
        user = torch.empty(B, dtype=torch.int32, device=self.device)
        hist = torch.empty((B, W), dtype=torch.int32, device=self.device)
        cand = torch.empty((B, 1 + K), dtype=torch.int32, device=self.device)
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        t = self.t
        self.step += 1
        _lib.check(self.lib.lstur_assemble_batch(
            B, W, K, p(t['sample_click']), p(idx), p(t['click_user']), p(t['stream_off']), p(t['stream_docs']),
            p(t['neg_off']), p(t['neg_docs']), ctypes.c_uint((self.seed * 1000003 + self.step) & 0xffffffff), p(user), p(hist),
            p(cand), ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return dict(user=user, hist_doc=hist, cand_doc=cand)

    def next_batch(self):
        assert self.n_samples >= self.B, 'fewer training samples than the batch size'
        return self.assemble(self._next_indices())
