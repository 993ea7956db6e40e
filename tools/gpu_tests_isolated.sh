#!/bin/bash
# Dev helper: run every GPU test function in its own process so that one CUDA fault
# (sticky context error) does not cascade into the following tests.
out=${1:-gpurun_out/isolated.log}
: > "$out"
for t in $(grep -o "^def test_[a-z0-9_]*" tests/test_gpu_parity.py | sed 's/def //'); do
  echo "=== $t" >> "$out"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$t" 2>&1 | tail -25 >> "$out"
done
grep -E "^=== |passed|failed" "$out"
