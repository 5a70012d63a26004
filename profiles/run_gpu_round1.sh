#!/bin/bash
# Round-1 GPU session: bench (both arms), ncu launch list, ncu --set full of the top kernels,
# GPU parity suite, side measurements.  Run under gpurun from the repo root; writes gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'^k_rx$|k_notch_apply|k_frontend|k_notch_guess' -s 8 -c 4 -f -o gpurun_out/prof_full \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
timeout 1100 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py --variant u8 --no-cpu > gpurun_out/bench_u8.json 2> gpurun_out/bench_u8.err
timeout 200 python bench.py --variant hs --no-cpu > gpurun_out/bench_hs.json 2> gpurun_out/bench_hs.err
tail -3 gpurun_out/pytest_gpu.log
head -c 600 gpurun_out/bench_n1.json
