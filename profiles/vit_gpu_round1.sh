#!/bin/bash
# Viterbi time-segment kernel: parity tests + side measurement.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "viterbi or fast_mode or shard or fastlock" > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -6 gpurun_out/pytest_gpu3.log
timeout 300 python bench.py --variant viterbi --no-cpu --steps 3 > gpurun_out/bench_viterbi.json 2> gpurun_out/bench_viterbi.err
head -c 3000 gpurun_out/bench_viterbi.json; tail -3 gpurun_out/bench_viterbi.err
