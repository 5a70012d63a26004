#!/bin/bash
# Second (last) GPU session of the end of round 2: parity suite, bench, launch list and ncu --set full captures of the
# kernels written since the previous profiles (Viterbi, seam plan, de-randomiser scan, deconvolution tiles, lock tracker).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/stages2.log; }
timeout 240 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
stamp "build rc=$?"
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -n 8 --durations=8 > gpurun_out/pytest_gpu.log 2>&1
stamp "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
stamp "bench rc=$?"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/bench_under_ncu.log 2>&1
stamp "launches rc=$?"
timeout 110 ncu --set full --clock-control none --import-source on \
    -k regex:'k_rx_plan|k_derand_tiles|k_derand_chain|k_derand_index|k_deconv_tiled|k_sync_track' -s 24 -c 9 -f -o gpurun_out/prof_ctl \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > gpurun_out/ncu_ctl.log 2>&1
stamp "ncu ctl rc=$?"
timeout 110 ncu --set full --clock-control none --import-source on -k regex:'k_viterbi' -s 3 -c 1 -f -o gpurun_out/prof_vit78 \
    python bench.py --variant viterbi78 --steps 1 --warmup 3 --no-cpu --packets 32768 > gpurun_out/ncu_vit78.log 2>&1
stamp "ncu vit78 rc=$?"
timeout 110 ncu --set full --clock-control none --import-source on -k regex:'k_viterbi' -s 3 -c 1 -f -o gpurun_out/prof_vit12 \
    python bench.py --variant viterbi --steps 1 --warmup 3 --no-cpu --packets 32768 > gpurun_out/ncu_vit12.log 2>&1
stamp "ncu vit12 rc=$?"
timeout 60 python bench.py --variant viterbi78 --steps 3 --cpu-sample-packets 512 > gpurun_out/bench_viterbi78.json 2> gpurun_out/bench_viterbi78.err
stamp "viterbi78 rc=$?"
timeout 60 python bench.py --variant viterbi --steps 3 --cpu-sample-packets 2048 > gpurun_out/bench_viterbi.json 2> gpurun_out/bench_viterbi.err
stamp "viterbi rc=$?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
stamp "smoke rc=$?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_viterbi78", "bench_viterbi"):
    try:
        b = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value=%.0f" % b["value"], "ms=%.2f" % b["ms_per_step"], "e2e=%.0f" % b["e2e"]["value"], b.get("seams"), b.get("ts_bit_exact_vs_reference"))
        if "kernel_ms_per_step" in b: print("   ", {k: round(v, 3) for k, v in b["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "failed:", e)
PY
