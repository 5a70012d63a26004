#!/bin/bash
# Tuning sweep (environment knobs of k_rx / k_notch_apply) + re-run of selected GPU tests.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "tx or viterbi or fast_mode or cnr or spectrum or exact_mode_every" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -4 gpurun_out/pytest_gpu2.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu --steps 5 2>gpurun_out/sweep_$label.err | python -c "
import json,sys
b=json.loads(sys.stdin.read()); k=b['kernel_ms_per_step']
print('$label', 'value=%.0f'%b['value'], 'ms=%.3f'%b['ms_per_step'], 'e2e=%.0f'%b['e2e']['value'], 'rx=%.3f'%k['rx'], 'compact=%.3f'%k['rx_compact'], 'notch=%.3f'%k['notch_apply'], 'guess=%.3f'%k['notch_guess'], 'tel=%.3f'%b['stage_wall_ms_per_step'].get('telemetry',0), 'seams', b['seams'])
" | tee -a gpurun_out/sweep.txt
}
run base LDVB_DUMMY=1
run spans75k LDVB_RX_SPANS=75000
run spans94k LDVB_RX_SPANS=94000
run spans114k LDVB_RX_SPANS=114000
run spans94k_c20 LDVB_RX_SPANS=94000 LDVB_RX_CARVEOUT=20
run spans94k_c50 LDVB_RX_SPANS=94000 LDVB_RX_CARVEOUT=50
run warm1 LDVB_NOTCH_WARM=1
run seg2 LDVB_NOTCH_SEG=2
