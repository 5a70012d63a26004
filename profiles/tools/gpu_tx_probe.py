"""Throughput of the device-resident transmit chain (context for DESIGN.md; not a bench line)."""
import sys, time
import torch
sys.path.insert(0, ".")
import leansdr_b200 as P

npk = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
tx = P.Transmitter(ratio="6/5", power="37.5", agc=True, max_packets=npk)
st = torch.cuda.current_stream()
tx.set_stream(st.cuda_stream)
ts = torch.empty(npk * 188, dtype=torch.uint8, device=dev)
cap = tx.max_samples(npk)
iq = torch.empty(2 * cap, dtype=torch.float32, device=dev)
for it in range(4):
    tx.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    tx.tsgen_device(0, npk, ts.data_ptr())
    n = tx.process_device(ts.data_ptr(), npk, iq.data_ptr(), cap)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"tx: {npk} packets -> {n} samples in {ms:.2f} ms = {n / ms / 1e3:.1f} MS/s")
