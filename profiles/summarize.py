#!/usr/bin/env python
"""Builds the committed profile summaries from what a gpurun session left in gpurun_out/:

  launches.csv  (ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 1 --no-cpu)
  bench_n1.json (the un-profiled run of the same commit)
  prof_full.ncu-rep (ncu --set full of the top kernels; read with `ncu -i ... --page raw --csv`)

Usage: python profiles/summarize.py [gpurun_out] [round tag, default r01]
Writes profiles/<tag>_launches.csv, <tag>_launch_list.md, <tag>_bench_n1.json (+ reference / variants when present),
<tag>_ncu_full.md and ncu_traffic.json."""
import csv, io, json, os, re, shutil, subprocess, sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "gpurun_out")
tag = sys.argv[2] if len(sys.argv) > 2 else "r01"

PROFILE_NAME = {"k_rx": "rx", "k_notch_fir": "notch_fir", "k_fir_edges": "fir_edges", "k_notch_apply": "notch_apply", "k_notch_guess": "notch_guess", "k_notch_detect": "notch_detect",
                "k_notch_verify": "notch_verify", "k_frontend": "frontend", "k_rx_stitch": "rx_stitch", "k_rx_plan": "rx_plan",
                "k_rx_compact": "rx_compact", "k_deconv_tiled": "deconv_carry", "k_deconv": "deconv_carry", "k_sync_track": "sync_track",
                "k_sync_flags": "sync_flags", "k_realign": "realign", "k_rs": "deint_rs", "k_derand_scan": "derand",
                "k_derand_out": "derand", "k_derand_tiles": "derand", "k_derand_chain": "derand", "k_derand_index": "derand",
                "k_rx_plan_local": "rx_plan", "k_rx_plan_apply": "rx_plan", "k_meas_power": "meas_power", "k_meas_ema": "meas_ema"}

def base(kn):
    m = re.search(r"(k_[a-z0-9_]+)", kn)
    return m.group(1) if m else kn

def launch_list():
    p = os.path.join(src, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    shutil.copy(p, os.path.join(HERE, f"{tag}_launches.csv"))
    tot = defaultdict(float); cnt = defaultdict(int); per = defaultdict(list); serial = []
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = base(r["Kernel Name"])
        if k.startswith("k_tx"):
            continue                                   # input synthesis, outside the timed region
        if k == "k_rx_serial":
            serial.append(float(r["Metric Value"]) / 1e6)   # cold-start AGC settling passes: once per stream, outside the timed steps
            continue
        tot[k] += float(r["Metric Value"]) / 1e6; cnt[k] += 1
        per[k].append((r["Grid Size"], float(r["Metric Value"]) / 1e6))
    # the device-resident steps (whole 128 M-sample batch): per kernel, the launches with the largest grid
    def gsize(g): return eval(g.replace("(", "[").replace(")", "]"))[0]
    full = {}
    for k, v in per.items():
        gmax = max(gsize(g) for g, _ in v)
        sel = [t for g, t in v if gsize(g) == gmax]
        full[k] = sum(sel) / len(sel)
    full_sum = sum(full.values())
    bench = json.loads([l for l in open(os.path.join(src, "bench_n1.json")) if l.startswith("{")][-1])
    live = bench["kernel_ms_per_step"]; step = bench["ms_per_step"]
    allms = sum(tot.values())
    out = [f"# {tag} -- ncu launch list of the bench command (per-kernel device time)", "",
           "Command (gpurun, 1 GPU): `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "
           "gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity` ( the list covers "
           f"the device-resident steps and the pipelined host-push steps: {sum(cnt.values())} launches of the receive path; the transmit-chain "
           "kernels that synthesise the input are left out).",
           "Times under ncu are cold-cache and serialised: compare SHARES with the live CUDA-event numbers of `bench.py` "
           f"(right-hand columns, un-profiled run of the same commit: value {bench['value']/1e3:.1f} GS/s, {step:.2f} ms/step).", "",
           "The list covers the fresh check pass, the warm-up and the timed steps of the device-resident loop and the pushes of the e2e leg "
           "(one staged piece per push since round 2, i.e. whole batches everywhere).  Left out of the shares: `k_rx_serial`, the cold-start "
           f"AGC settling pass, which runs once per STREAM ({len(serial)} launches here, {sum(serial)/max(len(serial),1):.2f} ms each: the two handles "
           "of the bench, one fresh pass + one stream each) and never inside the timed steps.", "",
           "| kernel | launches | total ms (ncu, all) | share (all) | ms per whole-batch launch (ncu) | share (whole-batch) | ms/step live (bench.py CUDA events) | share live |",
           "|---|---|---|---|---|---|---|---|"]
    live_sum = sum(live.values())
    seen = set()
    for k in sorted(tot, key=tot.get, reverse=True):
        pn = PROFILE_NAME.get(k)
        lv = live.get(pn) if pn and pn not in seen else None
        if pn: seen.add(pn)
        out.append(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100*tot[k]/allms:.1f} % | {full[k]:.3f} | {100*full[k]/full_sum:.1f} % | " +
                   (f"{lv:.3f} | {100*lv/live_sum:.1f} % |" if lv is not None else "(in the row of the same stage) |  |"))
    out += ["", f"Sum of live kernel times: {live_sum:.2f} ms of the {step:.2f} ms step (the rest is host-side planning between launches)."]
    open(os.path.join(HERE, f"{tag}_launch_list.md"), "w").write("\n".join(out) + "\n")

def copies():
    for f, dst in (("bench_n1.json", f"{tag}_bench_n1.json"), ("bench_ref_n1.json", f"{tag}_bench_reference_n1.json"),
                   ("bench_u8.json", f"{tag}_bench_variant_u8.json"), ("bench_hs.json", f"{tag}_bench_variant_hs.json"),
                   ("bench_viterbi.json", f"{tag}_bench_variant_viterbi.json"), ("bench_viterbi78.json", f"{tag}_bench_variant_viterbi78.json"),
                   ("sweep.txt", f"{tag}_sweep_rx_notch_knobs.txt")):
        p = os.path.join(src, f)
        if os.path.exists(p) and os.path.getsize(p):
            shutil.copy(p, os.path.join(HERE, dst))

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]

def ncu_full():
    rep = os.path.join(src, "prof_full.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, check=True).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    n = 128324096
    out = [f"# {tag} -- `ncu --set full --clock-control none --import-source on` of the heaviest kernels at the bench size", "",
           "Command (gpurun, 1 GPU, profiles/run_gpu_round2.sh): `ncu --set full --clock-control none --import-source on -k "
           "regex:'^k_rx$|^k_notch_fir$|^k_notch_guess$|^k_rx_compact$' -s 8 -c 4 python bench.py --steps 1 --warmup 3 --no-cpu` "
           f"({n} f32 samples per launch).  Per launch, under the profiler (cold, serialised): use shares and ratios, not absolutes.", ""]
    traffic = {"_source": "ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 --no-cpu (%d f32 samples per launch), "
                          "profiles/run_gpu_round2.sh; dram__bytes_read.sum + dram__bytes_write.sum per launch" % n,
               "samples_per_launch": n, "kernels": {}}
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    for r in rows[2:]:
        kn = r[ix["Kernel Name"]]
        out += [f"## {kn}", "", "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in ix:
                out.append(f"| {w} | {r[ix[w]]} | {units[ix[w]]} |")
        out.append("")
        b = base(kn)
        if b in PROFILE_NAME:
            rd = float(r[ix["dram__bytes_read.sum"]]) * mult.get(units[ix["dram__bytes_read.sum"]], 1)
            wr = float(r[ix["dram__bytes_write.sum"]]) * mult.get(units[ix["dram__bytes_write.sum"]], 1)
            traffic["kernels"][PROFILE_NAME[b]] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                                                   "bytes_per_sample": (rd + wr) / n,
                                                   "gpu_time_ms_under_ncu": float(r[ix["gpu__time_duration.sum"]])}
    open(os.path.join(HERE, f"{tag}_ncu_full.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)

if __name__ == "__main__":
    launch_list(); copies(); ncu_full()
    print("profiles updated from", src)
