#!/bin/bash
# Round-2 GPU session.  Run under gpurun from the repo root; writes gpurun_out/.
#   STAGES="tests ref bench launches full variants smoke"  (default: tests bench launches smoke)
#   FULLK=regex of the kernels of the ncu --set full pass
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
STAGES=${STAGES:-"tests bench launches smoke"}
FULLK=${FULLK:-'k_rx|k_notch_apply|k_frontend|k_notch_guess'}
has() { case " $STAGES " in *" $1 "*) return 0;; esac; return 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
if has tests; then
timeout ${TEST_TIMEOUT:-1100} python -m pytest tests -m gpu -q -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
fi
if has ref; then
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
fi
if has bench; then
timeout 400 python bench.py ${BENCH_ARGS} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
fi
if has launches; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu ${BENCH_ARGS} > gpurun_out/bench_under_ncu.log 2>&1
fi
if has full; then
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"$FULLK" -s ${FULL_SKIP:-8} -c ${FULL_COUNT:-4} -f -o gpurun_out/prof_full \
    python bench.py --steps 1 --warmup 3 --no-cpu ${BENCH_ARGS} > gpurun_out/bench_under_ncu_full.log 2>&1
fi
if has variants; then
for v in ${VARIANTS:-u8 hs viterbi viterbi78}; do
  extra=""
  case $v in viterbi) extra="--cpu-sample-packets 2048";; viterbi78) extra="--cpu-sample-packets 512";; esac
  timeout 400 python bench.py --variant $v --steps 3 $extra ${VARIANT_ARGS} > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
fi
if has smoke; then
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
fi
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref_n1", "bench_u8", "bench_hs", "bench_viterbi", "bench_viterbi78"):
    try:
        b = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value=%.0f" % b["value"], "ms=%.2f" % b["ms_per_step"], "e2e=%.0f" % b["e2e"]["value"], b.get("seams"), b.get("ts_bit_exact_vs_reference"))
        if "kernel_ms_per_step" in b: print("   ", {k: round(v, 3) for k, v in b["kernel_ms_per_step"].items()}); print("   ", b.get("fast_vs_exact"))
    except Exception as e:
        print(f, "failed:", e)
PY
