import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import leansdr_b200 as P
import bench
dev = torch.device("cuda", 0)
NB = 6
iq_all = bench.gen_vector_device(65536 * NB, dev, torch, P)
n = (iq_all.numel() // 2 // NB) // 4096 * 4096
rx = P.Receiver(anf=1, rx_mode=P.RX_FAST, max_batch=n, device=0, fmt="f32", resample=True)
cap = n // 900 + 64
ts_dev = torch.empty(cap * 188, dtype=torch.uint8, device=dev)
rx.set_stream(torch.cuda.current_stream().cuda_stream)
for mode in ("continuous", "reset"):
    rx.reset()
    for b in range(NB):
        if mode == "reset":
            rx.reset()
        torch.cuda.synchronize()
        rx.profile(True)
        t0 = time.perf_counter()
        k = rx.process_device(iq_all.data_ptr() + (b if mode == "continuous" else 0) * n * 8, n, ts_dev.data_ptr(), cap)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        prof = rx.get_profile(); rx.profile(False)
        wall = {kk[5:]: round(v["ms_total"], 2) for kk, v in prof.items() if kk.startswith("wall:")}
        kern = sum(v["ms_total"] for kk, v in prof.items() if not kk.startswith("wall:"))
        m = rx.meas()
        print(mode, b, "ms=%.2f kernels=%.2f" % (dt, kern), wall, "packets", k, "notch_rep", m["notch_repaired"], "seams_rep", m["seams_repaired"], "settle", m["settle_passes"], flush=True)
