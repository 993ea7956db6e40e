#!/bin/bash
# compute-sanitizer passes over the kernels of the final library at small shapes.
# usage: tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] ...   -> gpurun_out/r2_sanitizer_<tool>.log
set -u
mkdir -p gpurun_out
SEL='test_x3_ or tcgen05 or test_flag_matrix_train_grads_vs_oracle or test_train_dropout_vs_oracle or test_forward_vs_oracle_shapes or test_segreduce or test_adam_dense or test_loss_modes or test_shard_plan or test_fast_full_catalogue or test_embed_all_single_launch or test_mim_peaked'
for tool in "$@"; do
  log=gpurun_out/r2_sanitizer_$tool.log
  timeout 1500 compute-sanitizer --tool "$tool" --error-exitcode 0 --print-limit 20 \
      python -m pytest tests/test_gpu_tc.py tests/test_gpu_attn.py tests/test_gpu_parity.py tests/test_gpu_dp.py tests/test_gpu_catalogue.py \
      -q -x -k "$SEL" -p no:cacheprovider > "$log" 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|error" "$log" | tail -5
done
