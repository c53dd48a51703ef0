#!/usr/bin/env python
"""LSTUR training throughput on B200 (BASELINE.json metric) — contract in the task statement.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA through the C-ABI)
  python bench.py --impl reference --gpus N ...            reference arm: the CPU restatement of the
                                                           reference's Keras graph on the host cores

A "step" is one LSTUR training step (forward + backward + Keras-Adam) over one batch of B
impressions per rank; value = impressions/s over all ranks (weak scaling: B fixed per rank).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'lstur_train_impressions_per_sec'
UNIT = 'impressions/s'


def flops_per_impression(sh, trainable_emb=False):
    """SURVEY.md §8d / BASELINE.md §3: algorithmic FLOPs (no dedup)."""
    T = sh.W + 1 + sh.K
    tok = T * sh.L
    conv = tok * 2 * sh.k * sh.E * sh.F
    rest = tok * 4 * sh.F + T * 2 * sh.F * sh.U + sh.W * 2 * (sh.U * 3 * sh.U + sh.U * 3 * sh.U) + (1 + sh.K) * 2 * sh.U
    fwd = conv + rest
    return dict(conv_fwd=conv, fwd=fwd, train=3 * fwd if trainable_emb else fwd + conv + 2 * rest)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), 'fallback'


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/r01_ncu_full_summary.json); None if absent."""
    p = os.path.join(ROOT, 'profiles', 'r01_ncu_full_summary.json')
    try:
        for row in json.load(open(p)):
            if kernel_substr in row['kernel']:
                return (row['dram_read_GB'] + row['dram_write_GB']) * 1e9
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_workload(sh, n_batches, rank, B):
    from mnexp_b200 import synth
    tok, _, _ = synth.make_docs(sh.n_news, sh.L, sh.vocab)
    P = synth.make_weights(sh, arch=sh.arch)
    batches, pad_frac = synth.make_batches(sh, n_batches, seed=1236 + 1000 * rank, B=B)
    return tok, P, batches, pad_frac


def cpu_reference_run(sh, steps, warmup, sample_B, threads=None):
    """The reference's Keras graph restated in torch-CPU fp32 (oracle/lstur_torch.py), all host threads,
    on a bounded sample of the workload: `sample_B` rows per step, same tables, dropout on, dense Keras Adam."""
    import torch
    from oracle import lstur_torch as ot
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    tok, P, batches, _ = build_workload(sh, warmup + steps, 0, sample_B)
    ora = ot.LsturOracle(P, arch=sh.arch, dtype=torch.float32, lr=1e-3, dropout=0.2)
    times = []
    for i, b in enumerate(batches):
        t0 = time.perf_counter()
        ora.train_step(b['user'], tok[b['hist_doc']], tok[b['cand_doc']], training=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.sum(times))
    return dict(value=sample_B * len(times) / t, unit=UNIT, cores=threads, kind='port',
                sample='%d steps of %d impressions of workload %s (same tables/shapes, dropout 0.2, dense Keras-Adam), '
                       'torch-CPU fp32 restatement of the Keras graph' % (len(times), sample_B, sh.name),
                ms_per_step=1e3 * t / len(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='C3')
    ap.add_argument('--precision', default=os.environ.get('LSTUR_PRECISION', 'auto'))
    ap.add_argument('--batch', type=int, default=0, help='rows per rank (default: the workload batch size)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cpu-sample', type=int, default=64)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    from mnexp_b200 import synth
    sh = synth.SHAPES[args.workload]
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    B = args.batch or sh.B
    config = dict(workload='%s: LSTUR-%s train step, %d users / %d news / %d vocab, L%d W%d K%d E%d F%d U%d, B=%d per rank'
                           % (sh.name, 'ini' if sh.arch == 'igru' else 'con', sh.n_users, sh.n_news, sh.vocab, sh.L, sh.W,
                              sh.K, sh.E, sh.F, sh.U, B),
                  global_batch=B * world, parallelism='dp%d' % world, dropout=0.2, optimizer='Keras-Adam',
                  word_emb_trainable=False)

    if args.impl == 'reference':
        if rank != 0:
            return 0
        r = cpu_reference_run(sh, max(1, args.steps), min(args.warmup, 1), args.cpu_sample)
        line = dict(metric=METRIC, value=r['value'], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                    data='synthetic', impl='reference', config=config,
                    cpu_baseline=dict(value=r['value'], unit=UNIT, cores=r['cores'], kind=r['kind'], sample=r['sample']),
                    e2e=dict(value=r['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from mnexp_b200 import _lib
    from mnexp_b200.dist import DataParallel
    from mnexp_b200.engine import LsturEngine
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = _lib.load()
    precision = args.precision
    if precision == 'auto':
        precision = 'fp16_tc' if lib.lstur_conv_tc_available() else 'fp32'
    n_batches = 6
    tok, P, batches, pad_frac = build_workload(sh, n_batches, rank, B)
    eng = LsturEngine(P, B, sh.W, 1 + sh.K, sh.L, arch=sh.arch, doc_tokens=tok, dropout=0.2, lr=1e-3,
                      precision=precision, sparse_user_adam=True)
    dp = DataParallel(eng)
    dbs = [eng.to_device_batch(b) for b in batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(args.warmup):
        dp.train_step(dbs[i % n_batches])
    barrier()
    # ---- timed region: device-resident inputs, CUDA events, probe events around the dominant kernel
    probe_id = 1
    probes = [(eng.new_event(), eng.new_event()) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = lib.lstur_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        eng.set_probe(probe_id, *probes[i])
        dp.train_step(dbs[i % n_batches])
    ev1.record()
    barrier()
    launches = int(lib.lstur_launch_count() - l0)
    eng.set_probe(0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms[0])
    value = B * world * args.steps / (ms_total / 1e3)
    probe_ms = float(np.mean([eng.elapsed_ms(a, b) for a, b in probes]))
    loss_last = eng.loss()

    # ---- e2e: host (pinned) batches in, loss out, every step
    e2e = None
    if not args.no_e2e:
        hb = [{k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()} for b in batches]
        h2d = sum(int(t.numel() * t.element_size()) for t in hb[0].values())
        for i in range(2):
            float(dp.train_step(eng.to_device_batch(hb[i], non_blocking=True))[0])
        barrier()
        ev0.record()
        for i in range(args.steps):
            db = eng.to_device_batch(hb[i % n_batches], non_blocking=True)
            loss_host = float(dp.train_step(db)[0].item())
        ev1.record()
        barrier()
        ms2 = torch.tensor([ev0.elapsed_time(ev1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = dict(value=B * world * args.steps / (float(ms2[0]) / 1e3), unit=UNIT, h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=4, api='LsturEngine.train_step(doc-id batch from pinned host memory) -> loss.item()')

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    pk, pk_kind = peaks()
    fl = flops_per_impression(sh)
    tensor_peak = pk['bf16_tflops_sustained']
    conv_tflops = fl['conv_fwd'] * B / (probe_ms / 1e3) / 1e12
    roofline = dict(bound='tensor', kernel='title Conv1D forward (implicit GEMM), precision=%s' % precision,
                    achieved=conv_tflops, peak=tensor_peak, unit='TFLOP/s', frac=conv_tflops / tensor_peak,
                    traffic=ncu_traffic('news_conv_tc_fwd') if (B == sh.B and sh.name == 'C3') else None,
                    traffic_unit='bytes of DRAM traffic per launch (ncu --set full, profiles/r01_ncu_full_summary.json)',
                    peak_kind='%s bf16_tflops_sustained' % pk_kind, kernel_ms=probe_ms,
                    step_frac_of_train_roofline=(value / world) * fl['train'] / (tensor_peak * 1e12),
                    flops_per_impression_train=fl['train'])
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(sh, 6, 1, args.cpu_sample)
        cpu.pop('ms_per_step', None)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype={'bf16_tc': 'bf16', 'fp16_tc': 'f16'}.get(precision, 'f32'), data='synthetic',
                config=dict(config, precision=precision, l2='inputs larger than L2: %d distinct batches, >%d MB activations per step'
                                                          % (n_batches, eng.ws_bytes >> 20),
                            user_adam='row-sparse (documented deviation from dense Keras-Adam)', hist_pad_frac=round(pad_frac, 3)),
                roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=launches, clocks=clocks, loss=loss_last)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
