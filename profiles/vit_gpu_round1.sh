#!/bin/bash
# Viterbi time-segment kernel: parity tests + side measurements.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "viterbi or time_sharded_ts_bit_exact" > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -4 gpurun_out/pytest_gpu3.log
for v in viterbi viterbi78; do
  timeout 300 python bench.py --variant $v --no-cpu --steps 3 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/bench_$v.json")); k=b["kernel_ms_per_step"]
    print("$v", "value=%.0f"%b["value"], "ms=%.2f"%b["ms_per_step"], "e2e=%.0f"%b["e2e"]["value"], "viterbi_ms=%.2f"%k.get("viterbi",0), "rx=%.2f"%k.get("rx",0), b["seams"], "ts", b["ts_packets_per_step"], "launches", b["gpu_launches"])
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_$v.err").read()[-1500:])
PY
done
