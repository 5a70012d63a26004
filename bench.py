#!/usr/bin/env python
"""bench.py -- IQ MS/s of the leandvb DVB-S QPSK CR 1/2 receive path on B200.

One "step" = one pass of the whole chain (front-end FIR -> notch -> receiver ->
deconvolution -> sync -> de-interleave -> RS -> de-randomise) over one batch of
synthetic IQ (BASELINE.json configs[1]: 2.4 MS/s f32 IQ, 2 MS/s QPSK 1/2,
`leandvb --f32 --resample`, i.e. 1.2 samples/symbol and a 5-tap low-pass).

  value  device-resident: the batch is already in HBM when the timed region starts
         (ldvb_process_device), TS packets stay in HBM.
  e2e    the reference-facing call with HOST buffers: ldvb_push from pinned memory
         (H2D inside the timed region) + ldvb_pull of the TS bytes (D2H inside).
  --impl reference   the unmodified reference `leandvb` (oracle/_ref, built from
         /root/reference in the dev container) on the box's host cores.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IQ MS/s end-to-end leandvb DVB-S QPSK CR1/2; TS bytes bit-exact vs CPU"
REF_FLAGS = ["--f32", "-f", "2400e3", "--sr", "2000e3", "--cr", "1/2", "--standard", "DVB-S", "--resample"]


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def gen_vector(npackets: int) -> np.ndarray:
    from tests import vectors as V
    return V.ref_iq(npackets, fmt="f32")


def run_reference_cpu(raw: np.ndarray, replicas: int, repeats: int):
    """Times oracle/_ref/leandvb on `raw` (page-cached file), `replicas` processes at once.
    Returns (MS/s aggregate, TS bytes of one replica)."""
    from oracle import oracle as O
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, f"ldvb_bench_{os.getpid()}.cf32")
    raw.tofile(path)
    ts = None
    best = None
    try:
        for _ in range(repeats):
            t0 = time.perf_counter()
            procs = []
            for r in range(replicas):
                out = subprocess.PIPE if r == 0 else subprocess.DEVNULL
                procs.append(subprocess.Popen([O.ref_bin("leandvb"), *REF_FLAGS], stdin=open(path, "rb"),
                                              stdout=out, stderr=subprocess.DEVNULL))
            ts0 = procs[0].stdout.read()
            for p in procs:
                p.wait()
            dt = time.perf_counter() - t0
            ts = ts0
            best = dt if best is None else min(best, dt)
    finally:
        os.unlink(path)
    n = raw.size // 2
    return replicas * n / best / 1e6, ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--packets", type=int, default=65536, help="TS packets in the synthetic stream (1 packet ~ 1958 samples)")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--anf", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    workload = {"workload": "C2: leantsgen|leandvbtx -f 6/5 --power 37.5 --agc -> leandvb --f32 --resample "
                            "-f 2400e3 --sr 2000e3 --cr 1/2 (QPSK, 1.2 samples/symbol, 5-tap FIR, anf=%d)" % a.anf,
                "packets": a.packets}

    # ---------------------------------------------------------------- reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        ncores = os.cpu_count() or 1
        sample_pk = min(a.packets, 4096)           # bounded sample: ~8 Mi samples per replica
        raw = gen_vector(sample_pk)
        n = raw.size // 2
        for _ in range(a.warmup):
            run_reference_cpu(raw, ncores, 1)
        t0 = time.perf_counter()
        vals = [run_reference_cpu(raw, ncores, 1)[0] for _ in range(a.steps)]
        dt = time.perf_counter() - t0
        v = float(np.mean(vals))
        single, _ = run_reference_cpu(raw, 1, 1)
        line = {"metric": METRIC, "value": v, "unit": "MS/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {**workload, "sample_packets": sample_pk},
                "cpu_baseline": {"value": v, "unit": "MS/s", "cores": ncores, "kind": "reference",
                                 "sample": f"{ncores} concurrent single-threaded leandvb processes (the reference has no "
                                           f"threads), each on the same {n} samples from page cache; one process alone: "
                                           f"{single:.1f} MS/s"},
                "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # -------------------------------------------------------------------- B200 arm
    import torch
    import leansdr_b200 as P

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    raw = gen_vector(a.packets)
    n = raw.size // 2
    mode = P.RX_FAST if a.mode == "fast" else P.RX_EXACT
    rx = P.Receiver(fmt="f32", resample=True, anf=a.anf, rx_mode=mode, max_batch=n, device=local)
    stream = torch.cuda.current_stream()
    rx.set_stream(stream.cuda_stream)
    iq_dev = torch.from_numpy(raw).to(dev)
    cap = n // 1900 + 64
    ts_dev = torch.empty(cap * 188, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        rx.reset()
        return rx.process_device(iq_dev.data_ptr(), n, ts_dev.data_ptr(), cap)

    for _ in range(W):
        npk = step()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    rx.profile(True)
    l0 = rx.meas()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        npk = step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    prof = rx.get_profile()
    rx.profile(False)
    meas = rx.meas()
    launches = meas["kernel_launches"] - l0
    ts_gpu = ts_dev[: npk * 188].cpu().numpy().reshape(-1, 188)

    # ---- e2e: host buffers through push/pull
    pinned = torch.from_numpy(raw).pin_memory()
    for _ in range(2):
        rx.reset(); rx.push_ptr(pinned.data_ptr(), n); rx.pull_all()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(a.steps):
        rx.reset()
        rx.push_ptr(pinned.data_ptr(), n)
        d2h = rx.pull_all().nbytes
    f1.record(stream)
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
    clk = clocks.summary()

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    total = n * a.steps * world
    value = total / (ms * 1e-3) / 1e6
    e2e_value = total / (e2e_ms * 1e-3) / 1e6

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (per-launch CUDA-event time, algorithmic bytes)
    peaks, peak_kind = measured_peaks()
    sym = meas["symbols"]
    omega = 1.2
    alg_bytes = {                      # per launch, see DESIGN.md "Kernels"
        "frontend": n * (8 + 8),                      # cf32 in + cf32 out (FIR, D=1)
        "notch_apply": n * (8 + 8),
        "rx": n * 8 + sym * 4,                        # cf32 in + softsymbol out
        "rx_compact": sym * 8,
        "deconv_carry": sym * 4 + sym // 8,
        "deint_rs": (sym // 8) * 2,
    }
    wall = {k[5:]: v["ms_total"] / a.steps for k, v in prof.items() if k.startswith("wall:")}
    prof = {k: v for k, v in prof.items() if not k.startswith("wall:")}
    kern = {k: v["ms_total"] / max(v["launches"], 1) for k, v in prof.items()}
    per_step = {k: v["ms_total"] / a.steps for k, v in prof.items()}
    dom = max(per_step, key=per_step.get) if per_step else None
    roof = None
    if dom:
        ab = alg_bytes.get(dom, n * 8)
        ach = ab / (kern[dom] * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_kind,
                "ms_per_launch": kern[dom], "algorithmic_bytes_per_launch": ab,
                "share_of_step": per_step[dom] / (ms / a.steps)}
    fir = None
    if "frontend" in kern:
        ach = alg_bytes["frontend"] / (kern["frontend"] * 1e-3) / 1e9
        fir = {"kernel": "frontend(FIR)", "achieved": ach, "frac": ach / peaks["hbm_gbs"], "unit": "GB/s",
               "ms_per_launch": kern["frontend"]}

    # ---- CPU baseline: the unmodified reference on this box, same vector, in the same run
    cpu = None
    ts_match = None
    if not a.no_cpu:
        sample_pk = min(a.packets, 16384)
        sample = raw[: 2 * min(n, sample_pk * 1958)]
        v, ts_ref = run_reference_cpu(sample, 1, 3)
        ref_pk = np.frombuffer(ts_ref, dtype=np.uint8).reshape(-1, 188)
        k = min(len(ref_pk), len(ts_gpu))
        ts_match = bool(k > 0 and np.array_equal(ref_pk[:k], ts_gpu[:k]) and
                        (len(ts_gpu) >= len(ref_pk) if sample.size == raw.size else True))
        cpu = {"value": v, "unit": "MS/s", "cores": 1, "kind": "reference",
               "sample": f"oracle/_ref/leandvb {' '.join(REF_FLAGS)} on the first {sample.size // 2} samples of the same "
                         f"vector (best of 3, file in page cache); host has {os.cpu_count()} cores, the reference uses 1"}

    line = {"metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {**workload, "samples_per_step_per_gpu": n, "rx_mode": a.mode,
                       "l2": "input batch (%d MB) larger than L2, re-read every step" % (raw.nbytes >> 20),
                       "parallelism": "time spans inside one GPU; one independent stream per GPU"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": int(raw.nbytes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_fir": fir,
            "kernel_ms_per_step": per_step, "stage_wall_ms_per_step": wall,
            "cpu_baseline": cpu,
            "ts_packets_per_step": int(npk), "ts_bit_exact_vs_reference": ts_match,
            "seams": {"total": meas["seams_total"], "repaired": meas["seams_repaired"], "notch_repaired": meas["notch_repaired"]}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
