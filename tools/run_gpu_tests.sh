#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one file must not poison the others) and writes the
# logs under gpurun_out/.  Usage: tools/run_gpu_tests.sh [file ...]
mkdir -p gpurun_out
files="$@"
if [ -z "$files" ]; then files="$(ls tests/test_gpu_*.py) tests/test_ref_pinned.py"; fi
rc=0
for f in $files; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -x --timeout=600 > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $name rc=$r: $(tail -1 gpurun_out/$name.log)"
  if [ $r -ne 0 ]; then rc=1; grep -E "^(FAILED|ERROR)|Error|error|assert" gpurun_out/$name.log | head -12; fi
done
exit $rc
