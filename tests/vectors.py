"""Synthetic DVB-S IQ for tests, smoke and bench.

The generator is the reference's own transmitter chain, leantsgen | leandvbtx
[| leanchansim] (apps/leantsgen.cc, apps/leandvbtx.cc:79-197,
apps/leanchansim.cc:115-189), run from the binaries that `make -C oracle ref`
builds in the dev container into oracle/_ref/ (they travel to the GPU box; the
reference sources do not).  The committed fixtures under tests/golden/ cover
the case where those binaries are absent.
"""
from __future__ import annotations

import os
import subprocess

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def have_ref() -> bool:
    return all(os.path.exists(O.ref_bin(b)) for b in ("leantsgen", "leandvbtx", "leanchansim", "leandvb"))


def ts_packets(n: int, start: int = 0) -> np.ndarray:
    """leantsgen-style numbered packets (apps/leantsgen.cc:37-47)."""
    p = np.zeros((n, 188), np.uint8)
    t = np.arange(start, start + n, dtype=np.uint32)
    for i in range(0, 185, 4):
        p[:, i] = i
        p[:, i + 1] = (t >> 16) & 0xff
        p[:, i + 2] = (t >> 8) & 0xff
        p[:, i + 3] = t & 0xff
    p[:, 0] = 0x47
    return p


def ref_iq(npackets: int, ratio: str = "6/5", cr: str = "1/2", power: float = 37.5,
           noise_db: float | None = None, extra_chansim: list[str] | None = None,
           fmt: str = "f32", cst: str = "QPSK") -> np.ndarray:
    """leantsgen -c N | leandvbtx --const CST --cr CR -f RATIO --power P --agc [| leanchansim ...]."""
    ts = subprocess.run([O.ref_bin("leantsgen"), "-c", str(npackets)], stdout=subprocess.PIPE,
                        check=True).stdout
    iq = subprocess.run([O.ref_bin("leandvbtx"), "--const", cst, "--cr", cr, "-f", ratio, "--power", str(power), "--agc"],
                        input=ts, stdout=subprocess.PIPE, check=True).stdout
    if noise_db is not None or extra_chansim or fmt == "u8":
        args = [O.ref_bin("leanchansim"), "--if32"]
        if noise_db is not None:
            args += ["--awgn", str(noise_db), "--deterministic"]
        if extra_chansim:
            args += extra_chansim
        args += ["--ou8" if fmt == "u8" else "--of32"]
        iq = subprocess.run(args, input=iq, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                            check=True).stdout
    return np.frombuffer(iq, dtype=np.uint8 if fmt == "u8" else np.float32).copy()


def ref_iq_slice(first: int, count: int, margin_packets: int = 4000) -> np.ndarray:
    """f32 samples [first, first + count) of the endless stream that
    `leantsgen | leandvbtx --cr 1/2 -f 6/5 --power 37.5 --agc` would produce, without
    generating everything in front of them: the transmitter is fed numbered packets from a
    multiple of 40 (8-packet PRBS cycle, 5 packets = 9792 samples exactly) that lies
    `margin_packets` before the slice; its state (interleaver, filter, AGC) has converged to
    the bit-identical trajectory long before the slice starts (checked by bench.py: the halo a
    rank generates itself equals the one it receives from its neighbour)."""
    spp5 = 9792                                   # samples per 5 packets at 6/5 samples/symbol
    p0 = max(0, first * 5 // spp5 - margin_packets) // 40 * 40
    p1 = -(-(first + count) * 5 // spp5) + 64     # the transmitter keeps ~11 packets in flight
    ts = ts_packets(p1 - p0, p0).tobytes()
    iq = subprocess.run([O.ref_bin("leandvbtx"), "--cr", "1/2", "-f", "6/5", "--power", "37.5", "--agc"],
                        input=ts, stdout=subprocess.PIPE, check=True).stdout
    off = first - p0 // 5 * spp5
    a = np.frombuffer(iq, dtype=np.float32)
    if a.size < 2 * (off + count):
        raise RuntimeError("transmitter produced fewer samples than planned")
    return a[2 * off: 2 * (off + count)].copy()


def make_iq(npackets: int = 300, fmt: str = "f32", **kw) -> np.ndarray:
    """Reference transmitter when oracle/_ref is present, else the committed u8 fixture
    (converted with the reference's own cconverter arithmetic for f32 requests)."""
    if have_ref():
        return ref_iq(npackets, fmt=fmt, **kw)
    g = np.fromfile(os.path.join(GOLDEN, "c1_160.u8"), dtype=np.uint8)
    if fmt == "u8":
        return g
    return (g.astype(np.int32) - 128).astype(np.float32)


def ref_leandvb(raw: np.ndarray, flags: list[str]) -> np.ndarray:
    """Runs the unmodified reference receiver on `raw`; returns TS packets [n,188]."""
    out = subprocess.run([O.ref_bin("leandvb"), *flags], input=raw.tobytes(), stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, check=True).stdout
    return np.frombuffer(out, dtype=np.uint8).reshape(-1, 188).copy()
