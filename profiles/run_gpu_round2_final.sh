#!/bin/bash
# Last GPU session of round 2 (10 GPU-minutes were left): stages in order of importance, each under its own timeout,
# everything into gpurun_out/.  The GPU parity suite runs under pytest-xdist (8 workers share the GPU; the oracle side
# of every test is host work).  Run under gpurun from the repo root.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/stages.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
stamp "build"
timeout 240 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
stamp "build rc=$?"
timeout ${TEST_TIMEOUT:-280} python -m pytest tests -m gpu -q -p no:cacheprovider -n ${XDIST:-8} --durations=12 > gpurun_out/pytest_gpu.log 2>&1
stamp "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 90 python bench.py --variant viterbi78 --steps 3 --cpu-sample-packets 512 > gpurun_out/bench_viterbi78.json 2> gpurun_out/bench_viterbi78.err
stamp "viterbi78 rc=$?"
timeout 90 python bench.py --variant viterbi --steps 3 --cpu-sample-packets 2048 > gpurun_out/bench_viterbi.json 2> gpurun_out/bench_viterbi.err
stamp "viterbi rc=$?"
timeout 200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
stamp "bench rc=$?"
LDVB_VIT_WS=0 timeout 60 python bench.py --variant viterbi --steps 3 --no-cpu > gpurun_out/bench_viterbi_ws0.json 2> gpurun_out/bench_viterbi_ws0.err
stamp "viterbi ws0 rc=$?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
stamp "smoke rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/bench_under_ncu.log 2>&1
stamp "launches rc=$?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_viterbi78", "bench_viterbi", "bench_viterbi_ws0"):
    try:
        b = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value=%.0f" % b["value"], "ms=%.2f" % b["ms_per_step"], "e2e=%.0f" % b["e2e"]["value"], b.get("seams"), b.get("ts_bit_exact_vs_reference"))
        if "kernel_ms_per_step" in b: print("   ", {k: round(v, 3) for k, v in b["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "failed:", e)
PY
