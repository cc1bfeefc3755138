#!/bin/bash
# Run the GPU parity tests file by file (own process each, so a faulting kernel cannot poison
# the CUDA context of the rest); logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
for w in 0 1 2 3 4; do
  timeout 300 python -m pytest "tests/test_umma_selftest.py::test_umma_operand_forms[$w]" -q -x -p no:cacheprovider > gpurun_out/selftest_$w.log 2>&1
  echo "selftest[$w] exit $?"; tail -3 gpurun_out/selftest_$w.log
done
for f in test_rowwise_gpu test_rope_gpu test_ring_ops_gpu test_attention_gpu test_attention_coverage_gpu test_varlen_gpu test_llama_block_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 900 "$@" > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; grep -E "passed|failed|error" gpurun_out/$f.log | tail -3
  grep -E "^(FAILED|ERROR)" gpurun_out/$f.log | head -40
done
