#!/bin/bash
# Round-1 GPU session: GPU parity suite, bench (both arms), ncu launch list, [ncu --set full of the top kernels],
# side measurements.  Run under gpurun from the repo root; writes gpurun_out/.  FULL=1 adds the ncu --set full pass.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 1100 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
if [ -n "$FULL" ]; then
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'k_rx$|k_notch_apply|k_frontend|k_notch_guess' -s 8 -c 4 -f -o gpurun_out/prof_full \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
fi
for v in u8 hs viterbi viterbi78; do
  timeout 200 python bench.py --variant $v --no-cpu --steps 3 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref_n1", "bench_u8", "bench_hs", "bench_viterbi", "bench_viterbi78"):
    try:
        b = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value=%.0f" % b["value"], "ms=%.2f" % b["ms_per_step"], "e2e=%.0f" % b["e2e"]["value"], b.get("seams"), b.get("ts_bit_exact_vs_reference"))
    except Exception as e:
        print(f, "failed:", e)
PY
