#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): contract bench line, reference arm, ncu launch
# list of the same command, full ncu captures of the forward and backward main kernels.  $1 = tag.
cd "$(dirname "$0")/.."
TAG=${1:-r01c}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 600 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
echo "reference arm exit $?"; tail -c 300 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu1_$TAG.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_sm100_persist -s 3 -c 1 -f -o gpurun_out/prof_bwd_$TAG \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu2_$TAG.log 2>&1
echo "bwd capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_sm100 -s 3 -c 1 -f -o gpurun_out/prof_fwd_$TAG \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu3_$TAG.log 2>&1
echo "fwd capture exit $?"
ls -la gpurun_out/*$TAG*
