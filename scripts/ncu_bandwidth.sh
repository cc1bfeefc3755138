#!/bin/bash
# ncu --set full of every HBM-bound kernel (norms, softmax, RoPE, attention prep / post) at config C3
# shapes -> gpurun_out/prof_bw_$TAG.ncu-rep + a per-kernel summary (dram bytes vs algorithmic).  $1 = tag.
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
python scripts/ncu_bandwidth_driver.py > gpurun_out/ncu_bw_plain_$TAG.log 2>&1 || { echo "driver failed"; tail -5 gpurun_out/ncu_bw_plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on \
    -k regex:'rowwise_|reduce_partials|llama_rope|attn_bwd_prep|attn_bwd_post|pair_|dpair_|absmax_f32|split_f32|attn_bwd_f32_prep' -c 60 -f \
    -o gpurun_out/prof_bw_$TAG python scripts/ncu_bandwidth_driver.py > gpurun_out/ncu_bw_$TAG.log 2>&1
echo "bandwidth capture exit $?"
python scripts/ncu_summary.py gpurun_out/prof_bw_$TAG.ncu-rep gpurun_out/bw_ncu_full_$TAG.json > /dev/null
ls -la gpurun_out/*bw*$TAG*
