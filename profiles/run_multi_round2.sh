#!/bin/bash
# Round-2 multi-GPU session: bench.py on N GPUs of one box (time-sharded stream), library ring and python ring.
#   N=2 RINGS="lib py" bash profiles/run_multi_round2.sh        (under gpurun --gpus N)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${N:-2}
port=29600
for ring in ${RINGS:-lib py}; do
  port=$((port + 1))
  timeout ${RUN_TIMEOUT:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --ring $ring ${BENCH_ARGS} > gpurun_out/bench_n${N}_${ring}.json 2> gpurun_out/bench_n${N}_${ring}.err
  echo "ring=$ring rc=$?"
  tail -3 gpurun_out/bench_n${N}_${ring}.err | cut -c1-300
  python - "$N" "$ring" <<'PY'
import json, sys
n, ring = sys.argv[1], sys.argv[2]
try:
    b = json.loads([l for l in open(f"gpurun_out/bench_n{n}_{ring}.json") if l.startswith("{")][-1])
    print(f"N={n} ring={ring} value={b['value']:.0f} ms={b['ms_per_step']:.3f} e2e={b['e2e']['value']:.0f} ({b['e2e']['ms_per_step']:.2f} ms) "
          f"ts_ok={b['ts_equals_transmitted_packets_contiguous_over_ranks']} ref={b['ts_bit_exact_vs_reference']} halo={b['halo_received_equals_own_generation']}")
    for r in b["shard_timeline_ms_per_step"]:
        print("   ", {k: round(v, 3) for k, v in r.items()})
except Exception as e:
    print("failed:", e)
PY
done
