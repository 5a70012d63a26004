#!/usr/bin/env python
"""Summarises any `ncu --set full` report into a markdown table per captured launch (the metrics of summarize.py's WANT
list), plus the hottest source lines by warp stall samples when the report carries source counters.

Usage: python profiles/summarize_rep.py <report.ncu-rep> <out.md> "<command that produced it>" """
import csv, io, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv, argv = sys.argv[:1], sys.argv
from summarize import WANT  # noqa: E402

rep, out_md, cmd = argv[1], argv[2], argv[3] if len(argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, check=True).stdout.decode()
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
extra = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
         "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
         "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = [f"# {os.path.basename(out_md)[:-3]} -- `ncu --set full --clock-control none --import-source on`", "",
       f"Command (gpurun, 1 GPU): `{cmd}`.  Per launch, under the profiler (cold caches, serialised): use shares and ratios, not absolutes.", ""]
for r in rows[2:]:
    out += [f"## {r[ix['Kernel Name']]}", "", "| metric | value | unit |", "|---|---|---|"]
    for w in list(WANT) + extra:
        if w in ix:
            out.append(f"| {w} | {r[ix[w]]} | {units[ix[w]]} |")
    out.append("")
open(out_md, "w").write("\n".join(out) + "\n")
print("wrote", out_md, len(rows) - 2, "launches")
