"""Data-parallel LSTUR training: one process per GPU, batch sharded per rank.

The reference has no distribution at all (settings.py:97-99 drops the node-* options); this is the
exchange step a DP replica set needs: (1) one flat all-reduce of the dense-gradient arena,
(2) an all-gather of the per-sample user ids and user-embedding gradient rows, after which every
rank runs the same deterministic dedup + segment-sorted sum + row-sparse Adam, so replicas stay
bit-identical without broadcasting weights.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .engine import _ptr


def exchange(dense_grad, user_ids, d_u0, group=None):
    """Collective part only (works on gloo/CPU tensors too — exercised by tests/test_dist_cpu.py).

    dense_grad is summed in place across ranks; returns (all_ids (world*B,), all_rows (world*B, Ue))
    in rank-major order, or (None, None) when the model has no user embedding."""
    dist.all_reduce(dense_grad, op=dist.ReduceOp.SUM, group=group)
    if user_ids is None:
        return None, None
    world = dist.get_world_size(group)
    all_ids = torch.empty(world * user_ids.numel(), dtype=user_ids.dtype, device=user_ids.device)
    all_rows = torch.empty((world * d_u0.shape[0], d_u0.shape[1]), dtype=d_u0.dtype, device=d_u0.device)
    dist.all_gather_into_tensor(all_ids, user_ids.contiguous(), group=group)
    dist.all_gather_into_tensor(all_rows, d_u0.contiguous(), group=group)
    return all_ids, all_rows


def shard_batch(batch, rank, world):
    """Contiguous split of a global batch (dict of arrays with leading dim B_global) across ranks."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        assert n % world == 0, 'global batch must divide by world size'
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


MAX_SORT_KEYS = 16384      # lstur_sort_unique_i32 is a single-CTA bitonic sort (csrc/optim.cu: SORT_MAX)


class UserRowReducer:
    """Deterministic reduction of the exchanged (user id, d user_emb row) pairs of all ranks: the same dedup +
    segment-sorted sum on every rank (rank-major order), so the replicas stay bit-identical."""

    def __init__(self, engine, n):
        if n > MAX_SORT_KEYS:
            raise ValueError('data-parallel user-row dedup sorts world * B = %d keys; the limit is %d '
                             '(use fewer rows per rank or fewer ranks)' % (n, MAX_SORT_KEYS))
        dev = engine.device
        i32 = lambda k: torch.empty(k, dtype=torch.int32, device=dev)
        self.n = n
        self.sorted_pos, self.rows, self.seg, self.inv, self.nrows = i32(n), i32(n), i32(n + 1), i32(n), i32(1)
        self.grows = torch.empty((n, engine.Ue), dtype=torch.float32, device=dev)

    def reduce(self, engine, ids, rows):
        """-> the `user_rows` tuple LsturEngine.apply_adam takes"""
        n, st, L = ids.numel(), engine._stream(), engine.lib
        assert n == self.n
        _lib.check(L.lstur_sort_unique_i32(n, _ptr(ids), _ptr(self.sorted_pos), _ptr(self.rows), _ptr(self.seg),
                                           _ptr(self.inv), _ptr(self.nrows), st))
        _lib.check(L.lstur_segment_sum_rows(n, engine.Ue, _ptr(self.nrows), _ptr(self.seg), _ptr(self.sorted_pos),
                                            _ptr(rows), engine.Ue, _ptr(self.grows), st))
        return (n, self.rows, self.nrows, self.grows)


def init_process_group(local_rank, backend='nccl'):
    """One process per GPU; NCCL's internal stream runs at high priority so that the small exchange collectives get an
    SM as soon as one frees up under the (SM-filling) backward kernels they overlap with."""
    kw = {}
    if backend == 'nccl':
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            kw['pg_options'] = opts
        except Exception:       # noqa: BLE001  (older torch: default stream priority)
            pass
        kw['device_id'] = torch.device('cuda', local_rank)
    dist.init_process_group(backend, **kw)


class DataParallel:
    """Batch sharded per rank, weights replicated.  Per step:

      forward, backward on the main stream; lstur_backward records an event once every gradient except the title-encoder
      bucket (conv_w / conv_b / att_*: the first `head` floats of the arena) is final — about 2 ms before the end of the
      backward at C3.  A second (high-priority) stream waits for that event and runs, UNDER the attention backward and the
      conv weight gradient: the all-reduce of the arena tail, the all-gather of (user id, d user row), and the
      deterministic global dedup + segment-sorted sum.  Only the all-reduce of the title-encoder bucket (1.4 MB) — and
      the dense word-table gradient when the table is trainable — trails the backward.
    """

    def __init__(self, engine, group=None, overlap=True):
        self.eng, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        e = engine
        self.reducer = UserRowReducer(e, self.world * e.B) if (e.user_emb is not None and self.world > 1) else None
        self.overlap = bool(overlap) and self.world > 1
        if self.world > 1:
            self.head = int(e.lib.lstur_plan_dense_head_count(e.plan))
            if self.overlap:
                self.side = torch.cuda.Stream(device=e.device, priority=-1)
                self.ev_tail, self.ev_side = e.new_event(), e.new_event()
                _lib.check(e.lib.lstur_plan_set_event(e.plan, 1, self.ev_tail))

    def train_step(self, db):
        e = self.eng
        if self.world == 1:
            return e.train_step(db)
        e.step_seed += 1
        e.forward(db, training=True, seed=e.step_seed * self.world + self.rank)
        e.backward(db, grad_scale=1.0 / (e.B * self.world))          # mean over the GLOBAL batch
        has_user = e.user_emb is not None
        ids_in = db['user'] if has_user else None
        rows_in = e.view('d_u0').reshape(e.B, e.Ue) if has_user else None
        ur = None
        if self.overlap:
            side_h = ctypes.c_void_p(self.side.cuda_stream)
            _lib.check(e.lib.lstur_stream_wait_event(side_h, self.ev_tail))
            with torch.cuda.stream(self.side):
                ids, rows = exchange(e.dense_grad[self.head:], ids_in, rows_in, self.group)
                if has_user:
                    ur = self.reducer.reduce(e, ids, rows)
                _lib.check(e.lib.lstur_event_record(self.ev_side, side_h))
            dist.all_reduce(e.dense_grad[:self.head], op=dist.ReduceOp.SUM, group=self.group)
            if e.trainable_word_emb:
                dist.all_reduce(e.word_grad, op=dist.ReduceOp.SUM, group=self.group)
            _lib.check(e.lib.lstur_stream_wait_event(e._stream(), self.ev_side))
        else:
            ids, rows = exchange(e.dense_grad, ids_in, rows_in, self.group)
            if e.trainable_word_emb:
                dist.all_reduce(e.word_grad, op=dist.ReduceOp.SUM, group=self.group)
            if has_user:
                ur = self.reducer.reduce(e, ids, rows)
        e.apply_adam(user_rows=ur)
        return e.view('loss')


# ---- sharded decomposed inference (C4; SURVEY.md §8e "Inference"): news sharded by id, one all-gather of the document
# vectors, then users sharded by id with no further exchange ----------------------------------------------------------
def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n items owned by `rank`; every rank's slice has ceil(n / world) slots (the last
    ones may be short)."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per), per


def gather_rows(local_rows, n_total, group=None):
    """all-gather of equally sized per-rank row blocks (padded to ceil(n/world) rows) -> the first n_total rows.
    Collective part only: runs on gloo/CPU tensors too (tests/test_dist_cpu.py)."""
    world = dist.get_world_size(group)
    full = torch.empty((world * local_rows.shape[0],) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype,
                       device=local_rows.device)
    dist.all_gather_into_tensor(full, local_rows.contiguous(), group=group)
    return full[:n_total]


def sharded_doc_table(engine, group=None):
    """TestPipeline.test_doc_vec over all ranks: rank r encodes documents [lo, hi), the table is all-gathered (C4:
    130 k x U fp32 = 104 MB over NVLink) and row 0 (pad document) is zeroed like engine.build_doc_table()."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return engine.build_doc_table()
    n_docs = int(engine.doc_tokens.shape[0])
    lo, hi, per = shard_range(n_docs, rank, world)
    ids = torch.zeros(per, dtype=torch.int32, device=engine.device)
    ids[:hi - lo] = torch.arange(lo, hi, dtype=torch.int32, device=engine.device)
    table = gather_rows(engine.encode_docs(ids), n_docs, group)
    table[0].zero_()
    return table


def shard_users(n_users, group=None):
    """[lo, hi) of the users this rank scores against the shared document table (no exchange afterwards)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi, _ = shard_range(n_users, rank, world)
    return lo, hi
