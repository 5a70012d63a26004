"""Ad-hoc GPU bring-up probe (not a pytest): product vs oracle on a generated vector."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
import leansdr_b200 as P

R = O.ref_bin
def gen(npk, flags=("-f", "6/5", "--power", "37.5", "--agc")):
    ts = subprocess.run([R("leantsgen"), "-c", str(npk)], stdout=subprocess.PIPE).stdout
    iq = subprocess.run([R("leandvbtx"), *flags], input=ts, stdout=subprocess.PIPE).stdout
    return np.frombuffer(iq, dtype=np.float32).copy()

def cmp(name, a, b):
    a = np.ascontiguousarray(a).reshape(-1).view(np.uint8); b = np.ascontiguousarray(b).reshape(-1).view(np.uint8)
    n = min(a.size, b.size)
    eq = np.array_equal(a[:n], b[:n])
    first = -1 if eq else int(np.nonzero(a[:n] != b[:n])[0][0])
    print(f"  {name:11s} gpu={a.size:9d} oracle={b.size:9d} prefix_equal={eq} first_diff={first}", flush=True)
    return eq

def run(tag, raw, ocfg, pkw, batch=None):
    print(f"== {tag}", flush=True)
    t0 = time.time(); ref = O.Chain(ocfg).run(raw); t_or = time.time() - t0
    n = raw.size // 2
    rx = P.Receiver(keep_taps=1, max_batch=(batch or n), **pkw)
    taps = {k: [] for k in ("pp", "symbols", "bytes", "mpegbytes", "rspackets", "rtspackets")}
    ts = []
    t0 = time.time()
    step = batch or n
    for s in range(0, n, step):
        rx.push(raw[2 * s: 2 * min(n, s + step)])
        ts.append(rx.pull_all())
        for k in taps: taps[k].append(rx.tap(k))
    t_gpu = time.time() - t0
    ok = True
    for k in taps:
        ok &= cmp(k, np.concatenate(taps[k]), ref[k])
    tsg = np.concatenate(ts)
    ok &= cmp("ts", tsg, ref["ts"])
    print(f"  ts packets gpu={tsg.shape[0]} oracle={ref['ts'].shape[0]}  oracle {t_or:.2f}s gpu {t_gpu:.2f}s  meas={rx.meas()}", flush=True)
    rx.close()
    return ok

if __name__ == "__main__":
    raw = gen(int(sys.argv[1]) if len(sys.argv) > 1 else 600)
    allok = True
    allok &= run("f32 anf0 exact", raw, O.Config(fmt="f32", anf=0), dict(fmt="f32", anf=0))
    allok &= run("f32 anf0 resample exact", raw, O.Config(fmt="f32", anf=0, resample=True), dict(fmt="f32", anf=0, resample=True))
    allok &= run("f32 anf0 exact batched", raw, O.Config(fmt="f32", anf=0), dict(fmt="f32", anf=0), batch=100000)
    allok &= run("f32 anf1 resample exact", raw, O.Config(fmt="f32", resample=True), dict(fmt="f32", resample=True))
    allok &= run("f32 anf0 FAST", raw, O.Config(fmt="f32", anf=0), dict(fmt="f32", anf=0, rx_mode=1))
    print("ALL OK" if allok else "SOME FAILED")
