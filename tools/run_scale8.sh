#!/bin/bash
# One 8-GPU box: weak C3 (1024 rows / rank), strong C3 (1024 rows in total), weak C5, the 2-rank NCCL DP test.  Usage (GPU box): tools/run_scale8.sh
mkdir -p gpurun_out
run() {  # name, nproc, extra args
  local name=$1 n=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --no-extras --no-cpu-baseline "$@" > gpurun_out/scale_$name.log 2> gpurun_out/scale_$name.err
  tail -1 gpurun_out/scale_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$name', 'n', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'impr/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['config'].get('global_batch'))"
}
run weak_c3_n8 8
run weak_c3_n2 2
run strong_c3_n8 8 --scaling strong
run strong_c3_n2 2 --scaling strong
run weak_c5_n8 8 --workload C5
run weak_c3_train_n8 8 --trainable-emb
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q 2>&1 | tail -2
