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
# What bounds each kernel.  The HBM fraction is reported for all of them; for the serial
# recurrences it says how far the kernel is from being a streaming kernel, not how well it uses the memory system.
BOUND = {"frontend": "hbm", "notch_guess": "hbm", "notch_fir": "issue/latency", "rx": "latency", "notch_apply": "latency",
         "viterbi": "latency", "rx_compact": "hbm", "deconv_carry": "hbm"}
REF_FLAGS = ["--f32", "-f", "2400e3", "--sr", "2000e3", "--cr", "1/2", "--standard", "DVB-S", "--resample"]


def notch_guess_bytes(n, in_bytes=8, anf=1, sms=148):
    """Bytes k_notch_guess has to read per launch: the two 4096-sample blocks in front of every segment's warm-up
    (k_notch.cu: `(b + warm + 2 - block0) % seg_blocks > 1` returns at once); segment length as pipeline.cu chooses
    it for the fused kernel (16 rows per CTA, one wave of CTAs: two per SM with one notch slot, else one)."""
    nblocks = max(1, n // 4096)
    target = 16 * sms * (2 if anf == 1 else 1)
    seg = min(64, max(1, -(-nblocks // target)))
    return n * in_bytes * min(2, seg) // seg


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic(kernel: str, n_samples: int):
    """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    `ncu --set full` capture of this bench command (profiles/ncu_traffic.json), scaled by samples per launch
    when the batch differs from the captured one.  None if that kernel was not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return float(t["kernels"][kernel]["bytes_per_sample"]) * n_samples
    except Exception:
        return None


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


def gen_vector_device(npackets: int, dev, torch, P, ratio="6/5", fec="1/2"):
    """The same waveform synthesised in HBM by the B200 transmit chain (include/leandvb_b200_tx.h;
    bit-identical to leantsgen | leandvbtx, tests/test_gpu_tx.py): numbered TS packets ->
    randomizer, RS, interleaver, convolutional code, QPSK, RRC x6/5, AGC.  Returns a float32 device
    tensor of interleaved I/Q."""
    tx = P.Transmitter(ratio=ratio, fec=fec, power="37.5", agc=True, max_packets=npackets, device=dev.index or 0)
    ts = torch.empty(npackets * 188, dtype=torch.uint8, device=dev)
    tx.tsgen_device(0, npackets, ts.data_ptr())
    cap = tx.max_samples(npackets)
    iq = torch.empty(2 * cap, dtype=torch.float32, device=dev)
    n = tx.process_device(ts.data_ptr(), npackets, iq.data_ptr(), cap)
    torch.cuda.synchronize()
    tx.close()
    out = iq[: 2 * n].clone()
    del iq, ts
    torch.cuda.empty_cache()
    return out


def run_reference_cpu(raw: np.ndarray, replicas: int, repeats: int, flags=None):
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
                procs.append(subprocess.Popen([O.ref_bin("leandvb"), *(flags or REF_FLAGS)], stdin=open(path, "rb"),
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


def e2e_runnable(raw: np.ndarray, ref_flags, device: int):
    """The real drop-in, measured: oracle/_ref/leandvb_gpu -- the reference's own scheduler,
    pipebuf, file_reader and file_writer around gpu_dvbs_receiver (leansdr_b200/host/gpu_runnables.h) -- against the
    unmodified oracle/_ref/leandvb, both reading the same cf32 file (page cache) and writing TS to a file.  The GPU
    binary reports the wall clock of its scheduler loop (--gpu-timing): CUDA start-up and ldvb_create, which a
    long-running receiver pays once, are outside; the whole-process figure is given as well."""
    import re
    from oracle import oracle as O
    exe = O.ref_bin("leandvb_gpu")
    if not os.path.exists(exe):
        return None
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    n = raw.size // 2
    n1, n2 = min(n, 16 << 20), min(n, 96 << 20)
    base = os.path.join(d, f"ldvb_runnable_{os.getpid()}")
    paths = {n1: base + "_a.cf32", n2: base + "_b.cf32"}
    out = base + ".ts"
    res = {}
    try:
        for k, pth in paths.items():
            raw[: 2 * k].tofile(pth)

        def run(cmd, pth):
            t0 = time.perf_counter()
            with open(pth, "rb") as fi, open(out, "wb") as fo:
                pr = subprocess.run(cmd, stdin=fi, stdout=fo, stderr=subprocess.PIPE, check=True, timeout=600)
            dt = time.perf_counter() - t0
            m = re.search(rb"LDVB_TIMING samples=(\d+) seconds=([0-9.]+)", pr.stderr or b"")
            with open(out, "rb") as f:
                return dt, f.read(), (int(m.group(1)), float(m.group(2))) if m else None
        gpu_cmd = [exe, *ref_flags, "--gpu-batch", str(1 << 25), "--gpu-device", str(device), "--gpu-timing"]
        _, ts1, _ = run(gpu_cmd, paths[n1])                       # page cache, driver, first-touch
        t2, ts2, loop = run(gpu_cmd, paths[n2])
        tr, tsr, _ = run([O.ref_bin("leandvb"), *ref_flags], paths[n1])
        a1 = np.frombuffer(ts1, np.uint8).reshape(-1, 188); ar = np.frombuffer(tsr, np.uint8).reshape(-1, 188)
        k = min(len(a1), len(ar))
        res = {"value": (loop[0] / loop[1] / 1e6) if loop else None, "unit": "MS/s",
               "how": f"leandvb_gpu --gpu-batch {1 << 25} --gpu-timing < {n2}-sample cf32 file > file: samples / wall clock of the reference "
                      "scheduler's loop (file_reader -> gpu_dvbs_receiver -> file_writer); pipebuf page-locked once by the runnable "
                      "(ldvb_host_register), async_push",
               "whole_process_MSps": n2 / t2 / 1e6, "whole_process_seconds": t2,
               "reference": {"value": n1 / tr / 1e6, "unit": "MS/s", "how": f"leandvb < the {n1}-sample file > file, one process ({tr:.2f} s)"},
               "ts_identical_to_reference": bool(k > 0 and np.array_equal(a1[:k], ar[:k]) and abs(len(a1) - len(ar)) <= 1),
               "ts_packets": int(len(np.frombuffer(ts2, np.uint8)) // 188)}
    finally:
        for pth in list(paths.values()) + [out]:
            if os.path.exists(pth):
                os.unlink(pth)
    return res


def awgn(raw: np.ndarray, mer_db: float) -> np.ndarray:
    """The reference's own channel simulator (apps/leanchansim.cc: wgn_c + adder, --deterministic seed) on an f32
    vector: `leanchansim --if32 --awgn DB --deterministic --of32`, the generator the parity tests use.  DB is the
    noise standard deviation in dB (leanchansim.cc:248-249); the bench signal's RMS is ~68 = 36.7 dB, and its MER
    without any noise is ~14 dB already (linear interpolation at 1.2 samples per symbol, no matched filter)."""
    from oracle import oracle as O
    out = subprocess.run([O.ref_bin("leanchansim"), "--if32", "--awgn", str(mer_db), "--deterministic", "--of32"],
                         input=raw.tobytes(), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    return np.frombuffer(out, dtype=np.float32).copy()


def fast_vs_exact(P, raw: np.ndarray, rx_kw: dict, ref_flags, anf: int, device: int, sample_packets: int):
    """FAST-mode parity, MEASURED: on a bounded prefix of the bench vector, clean and with the
    reference's AWGN at 22 dB and 10 dB, the FAST receiver against the EXACT receiver (symbol by symbol: hard decision
    and soft cost) and both against the unmodified reference binary (TS packets)."""
    from tests import vectors as V
    n = min(raw.size // 2, sample_packets * 1958)
    base = np.ascontiguousarray(raw[: 2 * n])
    sent = V.ts_packets(sample_packets + 64)

    def run(x, mode, pieces=1):
        """-> symbols, TS, meas, seconds of the pushes.  pieces = 2: the vector arrives in two pushes, so that the
        second one starts from a carried (settled) loop state -- the steady state of a stream."""
        rx = P.Receiver(anf=anf, rx_mode=mode, max_batch=x.size // 2, device=device, keep_taps=1, **rx_kw)
        m = x.size // 2
        cuts = [0, m] if pieces == 1 else [0, (m // 2) // 4096 * 4096, m]
        syms, tss, dt = [], [], 0.0
        for a0, a1 in zip(cuts[:-1], cuts[1:]):
            t0 = time.perf_counter()
            rx.push(x[2 * a0: 2 * a1])
            dt += time.perf_counter() - t0
            tss.append(rx.pull_all())
            syms.append(rx.tap("symbols").view(np.uint32).copy())
        meas = rx.meas()
        rx.close()
        return np.concatenate(syms), np.concatenate(tss), meas, dt, [s.size for s in syms]

    def ids(ts):
        """numbered packets -> set of counters of the packets that are bit-correct transmitted packets"""
        if not len(ts):
            return set(), 0
        c = (ts[:, 5].astype(np.int64) << 16) | (ts[:, 6].astype(np.int64) << 8) | ts[:, 7]
        ok = (c < len(sent))
        good = ok & (ts == sent[np.minimum(c, len(sent) - 1)]).all(axis=1)
        return set(c[good].tolist()), int((~good).sum())

    out = {"sample_samples": int(n)}
    for name, db in (("clean", None), ("awgn_stddev_22dB", 22.0), ("awgn_stddev_25dB", 25.0)):
        x = base if db is None else awgn(base, db)
        se, te, me, dte, _ = run(x, P.RX_EXACT)
        sf, tf, mf, _, _ = run(x, P.RX_FAST)
        s2, t2, m2, _, sizes2 = run(x, P.RX_FAST, pieces=2)
        tr = np.frombuffer(subprocess.run([_ref_bin("leandvb"), *ref_flags], input=x.tobytes(), stdout=subprocess.PIPE,
                                          stderr=subprocess.DEVNULL, check=True).stdout, dtype=np.uint8).reshape(-1, 188)
        k = min(se.size, sf.size)
        ie, be = ids(te); i_f, bf = ids(tf); ir, br = ids(tr)
        kk = min(len(te), len(tr))
        out[name] = {
            "symbols": int(k), "symbol_count_diff": int(sf.size) - int(se.size),
            "hard_symbol_mismatch": int(((se[:k] >> 16) != (sf[:k] >> 16)).sum()),
            "cost_mismatch": int(((se[:k] & 0xffff) != (sf[:k] & 0xffff)).sum()),
            # soft cost = int16 difference of two squared distances (sdr.h:529-560); its scale is ~|cost| ~ 212 * 53
            "cost_abs_diff_mean": float(np.abs((se[:k] & 0xffff).astype(np.int16).astype(np.int32) - (sf[:k] & 0xffff).astype(np.int16).astype(np.int32)).mean()) if k else None,
            "cost_abs_diff_max": int(np.abs((se[:k] & 0xffff).astype(np.int16).astype(np.int32) - (sf[:k] & 0xffff).astype(np.int16).astype(np.int32)).max()) if k else None,
            "cost_abs_mean": float(np.abs((se[:k] & 0xffff).astype(np.int16).astype(np.int32)).mean()) if k else None,
            "ts_packets": {"reference": int(len(tr)), "exact": int(len(te)), "fast": int(len(tf))},
            "exact_ts_bit_identical_to_reference_prefix": bool(kk > 0 and np.array_equal(te[:kk], tr[:kk]) and abs(len(te) - len(tr)) <= 1),
            "ts_packets_differing": {"fast_vs_reference": len(i_f ^ ir), "fast_vs_exact": len(i_f ^ ie), "exact_vs_reference": len(ie ^ ir)},
            "wrong_packets_delivered": {"reference": br, "exact": be, "fast": bf},
            "seams": {"total": mf["seams_total"], "repaired": mf["seams_repaired"],
                      "accepted_with_mismatch": mf["seams_mismatch_accepted"], "settle_passes": mf["settle_passes"],
                      "max_dphase": mf["seam_max_dphase"], "max_dfreqw": mf["seam_max_dfreqw"], "max_dmu": mf["seam_max_dmu"]},
            "mer_db": me["mer"],
            # the same vector in two pushes: the symbols of the SECOND push come from spans that started from a carried,
            # settled AGC / frequency state (what every batch but the first of a stream sees)
            "steady_state": (lambda h0, kk2: {
                "symbols": int(kk2 - h0),
                "hard_symbol_mismatch": int(((se[h0:kk2] >> 16) != (s2[h0:kk2] >> 16)).sum()),
                "cost_mismatch": int(((se[h0:kk2] & 0xffff) != (s2[h0:kk2] & 0xffff)).sum()),
                "cost_abs_diff_mean": float(np.abs((se[h0:kk2] & 0xffff).astype(np.int16).astype(np.int32) -
                                                   (s2[h0:kk2] & 0xffff).astype(np.int16).astype(np.int32)).mean()) if kk2 > h0 else None,
                "ts_packets_differing_vs_reference": len(ids(t2)[0] ^ ir)})(sizes2[0], min(se.size, s2.size)),
            "exact_mode_MSps": (x.size // 2) / dte / 1e6,
        }
    return out


def _ref_bin(name):
    from oracle import oracle as O
    return O.ref_bin(name)


def bench_time_sharded(a, rank, world, local, W, workload, dist, torch, P):
    """N > 1: one stream of N * C samples, chunk k on rank k.  Per step the
    whole stream is demodulated once from the reset state: halo exchange, notch-bin chain,
    speculative front stages on all ranks concurrently, EDGE chain through the back stages."""
    from leansdr_b200 import shard as S
    from tests import vectors as V
    dev = torch.device("cuda", local)
    unit = 4096
    C = int(a.packets * 1958.4) // unit * unit
    # Every step of this arm restarts the stream at chunk 0 (each rank keeps ONE chunk resident, generated by the reference
    # transmitter on the host): the cold-start AGC settling pass -- serial, ~14 ms, once per STREAM -- would be paid by
    # rank 0 on every step, so it is switched off here (settle_chunks = -1, the documented knob).  TS is unaffected (QPSK
    # hard decisions do not depend on the AGC estimate; checked below); soft costs are ~13 % off the serial ones until the
    # estimate has settled (bench.py N = 1, fast_vs_exact).
    rx = P.Receiver(fmt="f32", resample=True, anf=a.anf, rx_mode=P.RX_FAST, max_batch=C + (1 << 17), device=local, settle_chunks=-1)
    stream = torch.cuda.current_stream()
    rx.set_stream(stream.cuda_stream)
    H = -(-rx.shard_min_halo() // unit) * unit
    chunks = S.plan_stream(world * C, world, unit, H)
    ch = chunks[rank]
    raw = V.ref_iq_slice(ch.abs_raw0, ch.n_halo + ch.n_chunk)
    buf = torch.from_numpy(raw).to(dev)                      # [halo | chunk], interleaved I/Q floats
    own_halo = buf[: 2 * ch.n_halo].clone()
    buf[: 2 * ch.n_halo].zero_()                             # the halo only ever arrives over NCCL
    cap = C // 1900 + 64
    ts_dev = torch.empty(cap * 188, dtype=torch.uint8, device=dev)
    engine = S.GpuEngine(rx, ts_dev.data_ptr(), cap)
    # The ring: inside the library (ldvb_ring_*: ncclSend / ncclRecv between neighbours, C ABI) unless --ring py
    # asks for the torch.distributed transport of leansdr_b200/shard.py (the same three shard calls underneath).
    use_lib = (a.ring == "lib")
    if use_lib:
        from leansdr_b200 import capi
        ids = torch.zeros(256, dtype=torch.uint8, device=dev)
        if rank == 0:
            try:
                ids.copy_(torch.frombuffer(bytearray(capi.ring_unique_id() + capi.ring_unique_id()), dtype=torch.uint8))
            except Exception as e:                   # no libnccl.so.2 for the library: zeros tell every rank
                sys.stderr.write(f"[bench] ldvb_ring_unique_id failed ({e}); using the torch.distributed ring\n")
        dist.broadcast(ids, 0)
        idb = ids.cpu().numpy().tobytes()
        ok = 1.0 if any(idb) else 0.0
        try:
            if ok:
                rx.ring_init(idb[:128], idb[128:], rank, world)
        except Exception as e:                       # all ranks fall back together (agreed below)
            sys.stderr.write(f"[bench] rank {rank}: ldvb_ring_init failed ({e}); using the torch.distributed ring\n")
            ok = 0.0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag[0]) < 1.0:
            use_lib = False
            try:
                rx.ring_destroy()
            except Exception:
                pass
    ring = None if use_lib else S.Ring(dist, dev)
    halo_send = buf[buf.numel() - 2 * ch.n_halo_next:] if ch.n_halo_next else None
    halo_recv = buf[: 2 * ch.n_halo] if ch.n_halo else None
    tl = {}

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def ring_flush():
        if use_lib:
            rx.ring_flush()
        else:
            ring.flush()

    def step(timeline=None, b=None):
        b = buf if b is None else b
        if use_lib:
            sh = rx.shard(b.data_ptr(), ch.abs_raw0, ch.n_halo, ch.n_chunk, ch.n_halo_next, ch.last)
            return rx.ring_round(sh, ts_dev.data_ptr(), cap)
        return S.run_round(engine, ring, ch, b.data_ptr(), b[b.numel() - 2 * ch.n_halo_next:] if ch.n_halo_next else None,
                           b[: 2 * ch.n_halo] if ch.n_halo else None, timeline)

    for _ in range(W):
        npk = step()
    ring_flush()
    torch.cuda.synchronize()
    halo_ok = bool(torch.equal(buf[: 2 * ch.n_halo], own_halo))
    if use_lib:
        rx.ring_stats(reset=True)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    rx.profile(True)
    l0 = rx.meas()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        npk = step(tl)
    ring_flush()
    e1.record(stream)
    barrier()
    ms = max(e0.elapsed_time(e1), 0.0)
    ms_host = (time.perf_counter() - t0) * 1e3
    if use_lib:
        tl = rx.ring_stats(reset=True)
    prof = rx.get_profile()
    rx.profile(False)
    meas = rx.meas()
    launches = meas["kernel_launches"] - l0
    ts_gpu = ts_dev[: npk * 188].cpu().numpy().reshape(-1, 188)

    # ---- e2e: the chunk comes from pinned host memory every step, the TS goes back to the host.  Two chunk buffers:
    # the samples of step s+1 arrive (side stream) while step s is being demodulated, like consecutive chunks of a
    # live stream; every step waits for its own samples, and the last TS copy lies inside the timed region.
    pinned = torch.from_numpy(raw[2 * ch.n_halo:]).pin_memory()
    ts_host = torch.empty(cap * 188, dtype=torch.uint8).pin_memory()
    bufs = [buf, torch.empty_like(buf)]
    side = torch.cuda.Stream()
    arrived = [torch.cuda.Event(), torch.cuda.Event()]

    def fetch(i):
        side.wait_stream(torch.cuda.current_stream())          # the buffer's previous step is done with it
        with torch.cuda.stream(side):
            bufs[i][2 * ch.n_halo:].copy_(pinned, non_blocking=True)
            arrived[i].record(side)

    def e2e_run(nsteps):
        k = 0
        fetch(0)
        for i in range(nsteps):
            cur = i & 1
            if i + 1 < nsteps:
                fetch(cur ^ 1)     # (its outgoing halo of two steps ago has left: ldvb_ring_round waits for that first)
            arrived[cur].synchronize()
            torch.cuda.current_stream().wait_event(arrived[cur])
            k = step(b=bufs[cur])
            ts_host[: k * 188].copy_(ts_dev[: k * 188], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return k
    e2e_run(2)
    ring_flush()
    barrier()
    t0 = time.perf_counter()
    k = e2e_run(a.steps)
    ring_flush()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.summary()

    # ---- correctness: decoded packets are the transmitted numbered packets, contiguous over ranks
    ctr = (ts_gpu[:, 1].astype(np.int64) << 16) | (ts_gpu[:, 2].astype(np.int64) << 8) | ts_gpu[:, 3]
    i0 = 3 if rank == 0 else 0                     # the stream opens with the interleaver's fill
    ok = len(ts_gpu) > i0 + 8 and np.array_equal(ts_gpu[i0:], V.ts_packets(len(ts_gpu) - i0, int(ctr[i0])))
    # rank 0 also decodes the first samples of the stream (= of its own chunk) with the unmodified reference
    ref_match = None
    if rank == 0 and not a.no_cpu:
        nref = min(C, 8 << 20)
        tsr = np.frombuffer(subprocess.run([_ref_bin("leandvb"), *REF_FLAGS], input=raw[: 2 * nref].tobytes(), stdout=subprocess.PIPE,
                                           stderr=subprocess.DEVNULL, check=True).stdout, dtype=np.uint8).reshape(-1, 188)
        kk = min(len(tsr), len(ts_gpu))
        ref_match = bool(kk > 100 and np.array_equal(tsr[:kk], ts_gpu[:kk]))
    mine = torch.tensor([float(ms), float(e2e_ms), float(ms_host), float(ctr[i0]) if len(ctr) > i0 else -1.0,
                         float(ctr[-1]) if len(ctr) else -1.0, float(ok), float(len(ts_gpu)), float(launches), float(halo_ok),
                         tl.get("early", 0.0), tl.get("front", 0.0), tl.get("wait_edge", 0.0), tl.get("back", 0.0),
                         float(meas["seams_repaired"]), float(meas["notch_repaired"])],
                        dtype=torch.float64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    if rank != 0:
        return
    allv = torch.stack(allv).cpu().numpy()
    ms = float(allv[:, 0].max()); e2e_ms = float(allv[:, 1].max())
    total = C * world * a.steps
    value = total / (ms * 1e-3) / 1e6
    e2e_value = total / (e2e_ms * 1e-3) / 1e6
    contiguous = all(allv[k, 3] == allv[k - 1, 4] + 1 for k in range(1, world))
    ts_ok = bool(allv[:, 5].all() and contiguous and allv[:, 8].all())
    peaks, peak_kind = measured_peaks()
    wall = {k[5:]: v["ms_total"] / a.steps for k, v in prof.items() if k.startswith("wall:")}
    prof = {k: v for k, v in prof.items() if not k.startswith("wall:")}
    kern = {k: v["ms_total"] / max(v["launches"], 1) for k, v in prof.items()}
    per_step = {k: v["ms_total"] / a.steps for k, v in prof.items()}
    nloc = ch.n_halo + ch.n_chunk
    alg_bytes = {"frontend": nloc * 16, "notch_apply": nloc * 16, "notch_fir": nloc * 16, "notch_guess": notch_guess_bytes(nloc, anf=a.anf) if "notch_fir" in kern else nloc * 8,   # (unfused: one-block segments, every block read)
                 "rx": nloc * 8 + int(nloc / 1.2) * 4}
    dom = max(per_step, key=per_step.get)
    ab = alg_bytes.get(dom, nloc * 8)
    ach = ab / (kern[dom] * 1e-3) / 1e9
    roof = {"kernel": dom, "bound": BOUND.get(dom, "hbm"), "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / peaks["hbm_gbs"], "traffic": ncu_traffic(dom, nloc), "peak_source": peak_kind, "ms_per_launch": kern[dom],
            "ns_per_sample": kern[dom] * 1e6 / nloc,
            "algorithmic_bytes_per_launch": ab, "share_of_step": per_step[dom] / (ms / a.steps), "rank": 0}
    fir = None
    if "frontend" in kern:
        achf = alg_bytes["frontend"] / (kern["frontend"] * 1e-3) / 1e9
        fir = {"kernel": "frontend(FIR)", "achieved": achf, "frac": achf / peaks["hbm_gbs"], "unit": "GB/s",
               "ms_per_launch": kern["frontend"]}
    line = {"metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload),
            "run": {"samples_per_step_per_gpu": C, "stream_samples_per_step": C * world, "rx_mode": "fast",
                    "stream": "every step restarts the stream at chunk 0; cold-start AGC settling pass off (settle_chunks = -1)",
                    "exchange": "per step and boundary one NCCL send/recv of %d halo samples (%d KB), 16 B of notch bins and a "
                                "%d-byte EDGE (carry state); front stages concurrent, back stages chained" % (H, H * 8 >> 10, engine.edge_size)},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": int(C * 8 * world),
                    "d2h_bytes_per_step": int(allv[:, 6].sum() * 188), "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": int(allv[:, 7].sum()),
            "roofline": roof, "roofline_fir": fir,
            "kernel_ms_per_step": per_step, "stage_wall_ms_per_step": wall,
            "cpu_baseline": None,
            "ts_packets_per_step": int(allv[:, 6].sum()),
            "ts_bit_exact_vs_reference": ref_match,
            "ts_bit_exact_vs_reference_how": "rank 0: oracle/_ref/leandvb on the first %d samples of the stream against the packets of chunk 0" % min(C, 8 << 20),
            "ring": "library (ldvb_ring_round: ncclSend/ncclRecv, two communicators)" if use_lib else "python (torch.distributed isend/recv)",
            "ts_equals_transmitted_packets_contiguous_over_ranks": ts_ok,
            "halo_received_equals_own_generation": bool(allv[:, 8].all()),
            "shard_timeline_ms_per_step": [{"rank": k, "early": allv[k, 9] / a.steps, "front": allv[k, 10] / a.steps,
                                            "wait_edge": allv[k, 11] / a.steps, "back": allv[k, 12] / a.steps,
                                            "device_ms": allv[k, 0] / a.steps} for k in range(world)],
            "seams": {"repaired": int(allv[:, 13].sum()), "notch_repaired": int(allv[:, 14].sum())}}
    print(json.dumps(line))


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
    ap.add_argument("--no-parity", action="store_true", help="skip the fast_vs_exact parity measurement")
    ap.add_argument("--cpu-sample-packets", type=int, default=16384, help="packets of the vector the cpu_baseline leg decodes with the "
                    "reference binary (the Viterbi variants need a small sample: the reference does 0.2 MS/s at 7/8)")
    ap.add_argument("--parity-packets", type=int, default=2048, help="packets of the bench vector the fast_vs_exact leg decodes "
                    "(three noise levels x EXACT + FAST + the reference binary)")
    ap.add_argument("--variant", default="f32", choices=["f32", "u8", "hs", "viterbi", "viterbi78"],
                    help="side measurements (N = 1): 'u8' = the same chain fed complex<u8> IQ (leandvb --u8), 'hs' = leandvb --u8 --hs "
                         "(fast_qpsk_receiver path), 'viterbi' = leandvb --f32 --resample --viterbi (viterbi_sync instead of deconvol_sync).  "
                         "The default 'f32' is BASELINE.json's configuration.")
    ap.add_argument("--cpu-gen", action="store_true", help="synthesise the IQ with the reference binaries on the host "
                    "instead of the B200 transmit chain (N = 1)")
    ap.add_argument("--ring", default="lib", choices=["lib", "py"], help="time-sharded transport: ldvb_ring_* inside the library "
                    "(NCCL from C) or leansdr_b200/shard.py over torch.distributed")
    ap.add_argument("--shard", default="time", choices=["time", "streams"],
                    help="N > 1: 'time' = ONE stream cut into N time chunks (halo + EDGE over NCCL), "
                         "'streams' = N independent streams (replicas)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    # `config` is the same dictionary in both arms (the driver compares them); what is specific to a run goes to `run`.
    workload = {"workload": "C2: leantsgen|leandvbtx -f 6/5 --power 37.5 --agc -> leandvb --f32 --resample "
                            "-f 2400e3 --sr 2000e3 --cr 1/2 (QPSK, 1.2 samples/symbol, 5-tap FIR, anf=%d)" % a.anf,
                "packets": a.packets, "variant": a.variant, "shard": a.shard if a.gpus > 1 else "none",
                "l2": "b200 arm: the per-GPU input batch (packets x 1958 samples x 8 B, 1 GB at the default size) is larger "
                      "than L2 and re-read every step; reference arm: host CPU",
                "parallelism": "b200 arm: N = 1 time spans inside one GPU, N > 1 ONE stream time-sharded over N GPUs (chunk k on "
                               "GPU k; halo, notch bins and EDGE carry state over NCCL send/recv); reference arm: one "
                               "single-threaded leandvb process per host core (the reference has no threads)"}

    ref_flags = list(REF_FLAGS)
    rx_kw = dict(fmt="f32", resample=True)
    tx_kw = {}
    if a.variant != "f32":
        if a.impl == "reference" or world > 1:
            raise SystemExit("--variant is a single-GPU side measurement of the b200 arm")
        if a.variant == "viterbi78":
            # BASELINE.json configs[2]: broadcast rate, 2 samples per symbol
            ref_flags = ["--f32", "-f", "55e6", "--sr", "27.5e6", "--cr", "7/8", "--standard", "DVB-S", "--viterbi"]
            rx_kw = dict(fmt="f32", Fs=55e6, Fm=27.5e6, fec="7/8", viterbi=True)
            tx_kw = dict(ratio="2", fec="7/8")
            workload["workload"] = ("SIDE MEASUREMENT (BASELINE.json configs[2]): leantsgen|leandvbtx --cr 7/8 -f 2 --power 37.5 --agc -> leandvb "
                                    + " ".join(ref_flags))
        elif a.variant == "viterbi":
            ref_flags = list(REF_FLAGS) + ["--viterbi"]
            rx_kw = dict(fmt="f32", resample=True, viterbi=True)
            workload["workload"] = "SIDE MEASUREMENT, not BASELINE.json's configuration: same f32 waveform -> leandvb " + " ".join(ref_flags)
        else:
            ref_flags = ["--u8", "-f", "2400e3", "--sr", "2000e3", "--cr", "1/2", "--standard", "DVB-S"] + (["--hs"] if a.variant == "hs" else ["--resample"])
            rx_kw = dict(fmt="u8", hs=True) if a.variant == "hs" else dict(fmt="u8", resample=True)
            workload["workload"] = "SIDE MEASUREMENT, not BASELINE.json's configuration: same waveform as complex<u8> IQ -> leandvb " + " ".join(ref_flags)

    # ---------------------------------------------------------------- reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        ncores = os.cpu_count() or 1
        sample_pk = min(a.packets, 16384)          # bounded sample: the same 32 M-sample prefix the b200 arm's cpu_baseline leg decodes
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
                "config": dict(workload), "run": {"sample_packets": sample_pk, "samples_per_process_per_step": n},
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

    if world > 1 and a.shard == "time":
        bench_time_sharded(a, rank, world, local, W, workload, dist, torch, P)
        dist.barrier()
        dist.destroy_process_group()
        return

    # ONE continuous stream of NB batches: the warm-up and the timed steps demodulate consecutive batches of it, the
    # way a receiver sees a signal (loop state, AGC, notch, sync carried from batch to batch).  A stream that restarts
    # every step would pay the cold start -- the serial AGC settling pass of the first FAST batch, ~14 ms -- every time.
    vector_check = None
    NB = min(W + a.steps, 40)                                # batches held in HBM (1 GB each at the default size)
    NBH = min(NB, 8)                                         # ... and in page-locked host memory for the e2e leg
    if a.cpu_gen:
        iq_all = torch.from_numpy(gen_vector(a.packets * NB)).to(dev)
        workload["synthesis"] = "oracle/_ref leantsgen | leandvbtx on the host"
    else:
        iq_all = gen_vector_device(a.packets * NB, dev, torch, P, **tx_kw)
        workload["synthesis"] = "B200 transmit chain (ldvbtx_*), bit-identical to leantsgen | leandvbtx"
    if a.variant in ("u8", "hs"):
        # leanchansim --ou8 = cconverter<f32,0,u8,128,1,1> (dsp.h:33-54): (u8)(128 + x), truncating
        iq_all = (iq_all + 128.0).to(torch.uint8)
    n = (iq_all.numel() // 2 // NB) // 4096 * 4096            # samples per batch
    pinned = torch.empty(2 * n * NBH, dtype=iq_all.dtype).pin_memory()
    pinned.copy_(iq_all[: 2 * n * NBH])
    torch.cuda.synchronize()
    raw = pinned[: 2 * n].numpy()                            # the first batch (what the CPU legs decode a prefix of)
    if rank == 0 and not a.cpu_gen and not tx_kw and a.variant == "f32":
        head = gen_vector(min(a.packets, 2048))              # the unmodified reference transmitter, same packets
        vector_check = bool(head.size > 1000000 and np.array_equal(head.view(np.uint32), raw[: head.size].view(np.uint32)))
    bps = 2 * iq_all.element_size()
    mode = P.RX_FAST if a.mode == "fast" else P.RX_EXACT
    rx = P.Receiver(anf=a.anf, rx_mode=mode, max_batch=n, device=local, **rx_kw)
    stream = torch.cuda.current_stream()
    rx.set_stream(stream.cuda_stream)
    cap = n // 900 + 64
    ts_dev = torch.empty(cap * 188, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the packets of the first batch of a fresh stream: what the reference binary and the host path are compared with
    rx.reset()
    npk = rx.process_device(iq_all.data_ptr(), n, ts_dev.data_ptr(), cap)
    ts_gpu = ts_dev[: npk * 188].cpu().numpy().reshape(-1, 188)

    pos = [0]

    ts_all = torch.empty(a.steps * cap * 188, dtype=torch.uint8, device=dev)      # the packets of every timed step

    def step(slot=None):
        b = pos[0] % NB
        if b == 0:
            rx.reset()                                       # (start of the stream; again only if steps + warmup > 40)
        pos[0] += 1
        dst = ts_dev.data_ptr() if slot is None else ts_all.data_ptr() + slot * cap * 188
        return rx.process_device(iq_all.data_ptr() + b * n * bps, n, dst, cap)

    for _ in range(W):
        step()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    rx.profile(True)
    m0 = rx.meas()
    l0 = m0["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ks = []
    barrier()
    e0.record(stream)
    for i in range(a.steps):
        ks.append(step(i))
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    prof = rx.get_profile()
    rx.profile(False)
    meas = rx.meas()
    launches = meas["kernel_launches"] - l0
    # the timed steps decoded consecutive numbered packets of the transmitted stream, bit for bit
    from tests import vectors as V
    ts_stream = np.concatenate([ts_all[i * cap * 188: i * cap * 188 + k * 188].cpu().numpy() for i, k in enumerate(ks)]).reshape(-1, 188)
    ctr = (ts_stream[:, 5].astype(np.int64) << 16) | (ts_stream[:, 6].astype(np.int64) << 8) | ts_stream[:, 7]
    stream_ok = bool(len(ts_stream) > 100 and np.array_equal(ts_stream, V.ts_packets(len(ts_stream), int(ctr[0]))))
    npk_step = len(ts_stream) // a.steps

    # ---- e2e: host buffers through push/pull, the way the runnable drives the handle (gpu_runnables.h): consecutive
    # batches of the same stream are pushed back to back from page-locked memory (async_push: a push returns when the
    # samples have left the host buffer, the chain of a batch overlaps the copy of the next one), packets are pulled as
    # they complete, and the final flush + pull lie inside the timed region.
    rx2 = P.Receiver(anf=a.anf, rx_mode=mode, max_batch=n, device=local, async_push=True, **rx_kw)
    ts_host = torch.empty(cap * 188, dtype=torch.uint8)       # the caller's TS buffer (ldvb_pull copies into it)
    # the host path gives the packets of the device-resident path (one fresh batch, outside the timed region)
    rx2.push_ptr(pinned.data_ptr(), n); rx2.flush()
    d2h = rx2.pull_ptr(ts_host.data_ptr(), cap) * 188
    e2e_ts_ok = bool(d2h == npk * 188 and np.array_equal(ts_host[:d2h].numpy().reshape(-1, 188), ts_gpu))
    rx2.reset()

    def drain():
        k = 0
        while True:
            got = rx2.pull_ptr(ts_host.data_ptr(), cap)
            if not got:
                return k
            k += got
    epos = [0]

    def e2e_push():
        b = epos[0] % NBH
        if b == 0 and epos[0]:
            rx2.reset()                                      # the host copy holds NBH batches: the stream restarts (a cold start)
        epos[0] += 1
        rx2.push_ptr(pinned.data_ptr() + b * n * bps, n)
    for _ in range(2):
        e2e_push(); drain()
    rx2.flush(); drain()
    barrier()
    # Producer / consumer: this thread pushes, a second host thread pulls (ldvb_pull may run next to ldvb_push on an
    # async_push handle: it only touches the packet queue).  The copy engine then never waits for a pull.
    pulled = [0]
    done = threading.Event()

    def puller():
        while True:
            got = rx2.pull_ptr(ts_host.data_ptr(), cap)
            pulled[0] += got
            if not got:
                if done.is_set():
                    return
                time.sleep(0.0002)
    t0 = time.perf_counter()
    th = threading.Thread(target=puller)
    th.start()
    for _ in range(a.steps):
        e2e_push()
    rx2.flush()
    done.set()
    th.join()
    e2e_packets = pulled[0] + drain()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.summary()
    d2h = e2e_packets * 188 // a.steps
    e2e_seams = rx2.meas()
    rx2.close()
    npk = npk_step

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
    sym = (meas["symbols"] - m0["symbols"]) // a.steps      # per step (the counters run over the whole stream)
    omega = 1.2
    alg_bytes = {                      # per launch, see DESIGN.md "Kernels"
        "frontend": n * ((2 if a.variant in ("u8", "hs") else 8) + 8),   # IQ in + cf32 out (FIR, D=1)
        "notch_apply": n * ((2 if a.variant in ("u8", "hs") else 8) + 8),
        "notch_fir": n * ((2 if a.variant in ("u8", "hs") else 8) + 8),    # IQ in + preprocessed cf32 out (notch + FIR fused)
        "notch_guess": (notch_guess_bytes(n, 2 if a.variant in ("u8", "hs") else 8, anf=a.anf) if "notch_fir" in prof
                        else n * (2 if a.variant in ("u8", "hs") else 8)),   # 2 blocks per segment, not the stream
        "viterbi": sym * 4 + sym // 8,
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
        traffic = ncu_traffic(dom, n) if a.variant == "f32" else None
        roof = {"kernel": dom, "bound": BOUND.get(dom, "hbm"), "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                "ms_per_launch": kern[dom], "ns_per_sample": kern[dom] * 1e6 / n, "algorithmic_bytes_per_launch": ab,
                "share_of_step": per_step[dom] / (ms / a.steps),
                "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram read + write per launch)" if traffic else None,
                "dram_frac": (traffic / (kern[dom] * 1e-3) / 1e9 / peaks["hbm_gbs"]) if traffic else None}
    # north_star's headline fraction: the FIR stage against the HBM roofline.  With auto_notch in front (anf >= 1, the default)
    # the low-pass runs on the store path of the notch kernel, so the stage IS that fused kernel; without it, k_frontend.
    fir = None
    if "frontend" in kern:
        ach = alg_bytes["frontend"] / (kern["frontend"] * 1e-3) / 1e9
        fir = {"kernel": "frontend(FIR)", "achieved": ach, "frac": ach / peaks["hbm_gbs"], "unit": "GB/s",
               "ms_per_launch": kern["frontend"]}
    elif "notch_fir" in kern:
        ach = alg_bytes["notch_fir"] / (kern["notch_fir"] * 1e-3) / 1e9
        tr = ncu_traffic("notch_fir", n) if a.variant == "f32" else None
        fir = {"kernel": "notch_fir (auto_notch + fir_filter fused: the notched stream never reaches HBM)", "bound": BOUND["notch_fir"],
               "achieved": ach, "frac": ach / peaks["hbm_gbs"], "unit": "GB/s", "ms_per_launch": kern["notch_fir"], "traffic": tr,
               "note": "k_frontend alone (anf = 0, or LDVB_NOTCH_FUSE=0) streams at 0.98 of the copy peak (profiles/r02_*)"}

    # every kernel that was captured with ncu --set full: algorithmic and DRAM-traffic bandwidth against the peak
    roof_all = []
    for k in ("frontend", "notch_guess", "notch_apply", "notch_fir", "rx", "rx_compact"):
        if k in kern and a.variant == "f32":
            tr = ncu_traffic(k, n)
            ab = alg_bytes.get(k, n * 8)
            roof_all.append({"kernel": k, "bound": BOUND.get(k, "hbm"), "ms_per_launch": kern[k], "ns_per_sample": kern[k] * 1e6 / n,
                             "achieved": ab / (kern[k] * 1e-3) / 1e9,
                             "frac": ab / (kern[k] * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": tr,
                             "dram_frac": (tr / (kern[k] * 1e-3) / 1e9 / peaks["hbm_gbs"]) if tr else None})

    # ---- CPU baseline: the unmodified reference on this box, same vector, in the same run
    cpu = None
    ts_match = None
    if not a.no_cpu:
        sample_pk = min(a.packets, a.cpu_sample_packets)
        sample = raw[: 2 * min(n, sample_pk * 1958)]
        try:
            v, ts_ref = run_reference_cpu(sample, 1, 3, ref_flags)
        except Exception as e:
            sys.stderr.write(f"[bench] cpu_baseline leg failed: {e!r}\n")
            v, ts_ref = float("nan"), b""
        ref_pk = np.frombuffer(ts_ref, dtype=np.uint8).reshape(-1, 188)
        k = min(len(ref_pk), len(ts_gpu))
        ts_match = bool(k > 0 and np.array_equal(ref_pk[:k], ts_gpu[:k]) and
                        (len(ts_gpu) >= len(ref_pk) if sample.size == raw.size else True))
        cpu = {"value": v, "unit": "MS/s", "cores": 1, "kind": "reference",
               "sample": f"oracle/_ref/leandvb {' '.join(ref_flags)} on the first {sample.size // 2} samples of the same "
                         f"vector (best of 3, file in page cache); host has {os.cpu_count()} cores, the reference uses 1"}

    runnable = None
    if not a.no_cpu and a.variant == "f32" and a.mode == "fast":
        try:
            runnable = e2e_runnable(raw, ref_flags, local)
        except Exception as e:                                  # never lose the bench line to the side measurement
            runnable = {"error": repr(e)[:300]}

    parity = None
    if not a.no_cpu and not a.no_parity and a.variant == "f32" and a.mode == "fast":
        del iq_all
        torch.cuda.empty_cache()
        try:
            parity = fast_vs_exact(P, raw, rx_kw, ref_flags, a.anf, local, min(a.packets, a.parity_packets))
        except Exception as e:                                  # never lose the bench line to a side measurement
            parity = {"error": repr(e)[:300]}

    line = {"metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if a.variant != "hs" else "int (u8/u16 angles, 64-bit PLL)", "data": "synthetic",
            "config": {k: v for k, v in workload.items() if k != "synthesis"},
            "run": {"samples_per_step_per_gpu": n, "rx_mode": a.mode, "input_bytes_per_step": int(raw.nbytes),
                    "synthesis": workload.get("synthesis"),
                    "stream": "one continuous stream of %d batches in HBM; warm-up and timed steps take consecutive batches (no reset in between); the e2e leg streams the first %d of them from page-locked host memory and restarts the stream when it runs out" % (NB, NBH)},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": int(raw.nbytes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_fir": fir, "roofline_kernels": roof_all,
            "kernel_ms_per_step": per_step, "stage_wall_ms_per_step": wall,
            "cpu_baseline": cpu,
            "ts_packets_per_step": int(npk), "ts_bit_exact_vs_reference": ts_match,
            "fast_vs_exact": parity,
            "e2e_ts_equals_device_resident_ts": e2e_ts_ok,
            "timed_steps_ts_equals_transmitted_packets_contiguous": stream_ok,
            "e2e_runnable": runnable,
            "e2e_mode": {"how": "streamed: consecutive batches of one continuous stream through ldvb_push (async_push, page-locked source) "
                                "on one host thread and ldvb_pull on a second one; final ldvb_flush and the last pull inside the timed region",
                         "packets": int(e2e_packets), "seams_repaired": int(e2e_seams["seams_repaired"]),
                         "seams_total": int(e2e_seams["seams_total"])},
            "vector_equals_reference_transmitter_prefix": vector_check,
            "seams": {"total": meas["seams_total"], "repaired": meas["seams_repaired"], "notch_repaired": meas["notch_repaired"],
                      "viterbi_segments": meas["vit_segments"], "viterbi_repaired": meas["vit_repaired"]}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
