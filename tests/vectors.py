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
           fmt: str = "f32") -> np.ndarray:
    """leantsgen -c N | leandvbtx --cr CR -f RATIO --power P --agc [| leanchansim ...]."""
    ts = subprocess.run([O.ref_bin("leantsgen"), "-c", str(npackets)], stdout=subprocess.PIPE,
                        check=True).stdout
    iq = subprocess.run([O.ref_bin("leandvbtx"), "--cr", cr, "-f", ratio, "--power", str(power), "--agc"],
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
