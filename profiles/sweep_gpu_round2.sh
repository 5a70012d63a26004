#!/bin/bash
# Round-2 knob sweep: environment knobs of the kernels, one short bench run each (no CPU legs).
#   SWEEP="VAR=v1,v2,... [VAR2=...]"   each variable is swept on its own, the others at their defaults
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
out=gpurun_out/sweep.txt
: > $out
for spec in ${SWEEP:-"LDVB_RX_WARPS=12,16,19,20"}; do
  var=${spec%%=*}; vals=${spec#*=}
  for v in ${vals//,/ }; do
    env $var=$v timeout 300 python bench.py --no-cpu --no-parity --steps 3 ${BENCH_ARGS} > gpurun_out/sweep_run.json 2> gpurun_out/sweep_run.err
    python - "$var" "$v" >> $out <<'PY'
import json, sys
try:
    b = json.load(open("gpurun_out/sweep_run.json"))
    k = b["kernel_ms_per_step"]
    print(sys.argv[1], sys.argv[2], "value=%.0f ms=%.3f e2e=%.0f" % (b["value"], b["ms_per_step"], b["e2e"]["value"]),
          " ".join("%s=%.3f" % (n, k[n]) for n in ("rx", "notch_apply", "notch_guess", "notch_fir", "frontend", "rx_compact") if n in k), b.get("seams"))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed:", e, open("gpurun_out/sweep_run.err").read()[-400:])
PY
  done
done
cat $out
