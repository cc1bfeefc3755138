#!/bin/bash
# One multi-GPU box (gpurun --gpus 8): ring parity on every GPU, config C5 through the C-ABI ring, the
# copy-only diagnosis of the host-buffer path at 1 / 4 / 8 ranks, and the strong-scaling record of config C2.
cd "$(dirname "$0")/.."
TAG=${1:-r02}
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_ring_ops_gpu.py tests/test_ring.py -m gpu -q -p no:cacheprovider \
  -k "p2p_ring_on_all_gpus or nccl_two_gpus" 2>&1 | tail -4 | tee gpurun_out/ring_${N}gpu_tests_$TAG.log
timeout 600 python scripts/perf_ring_p2p.py 131072 4 2>&1 | tee gpurun_out/ring_p2p_c5_$TAG.log
timeout 300 python scripts/perf_copy_only.py 2>&1 | tee gpurun_out/copy_only_n1_$TAG.log
for n in 4 $N; do
  timeout 300 $TR --nproc-per-node $n --master-port 2951$n scripts/perf_copy_only.py 2>&1 | grep -E "^#|GB/s" | tee gpurun_out/copy_only_n${n}_$TAG.log
done
for n in 2 4 $N; do
  timeout 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 --scaling strong \
    --no-cpu --no-secondary 2> gpurun_out/bench_strong_n${n}_$TAG.err | grep '^{' > gpurun_out/bench_strong_n${n}_$TAG.json
  echo "strong n=$n exit $?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_strong_n${n}_$TAG.json").read().strip().splitlines()[-1])
print("strong N=${n}: value", round(d["value"],1), "TFLOP/s, ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "ms", round(d["e2e"]["ms_per_step"],2))
PY
done
