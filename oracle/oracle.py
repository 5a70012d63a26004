"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/liboracle.so (the plain-C CPU restatement of the
leandvb DVB-S receive path, oracle/dvbs_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.

`Chain` wires the stages in the order of the reference front end
(/root/reference/src/apps/leandvb.cc:204-596) and runs them stage by stage on
whole arrays ("one-shot" schedule: every stage sees the complete output of the
previous one, which is the limit of the reference scheduler for large
--buf-factor; steady-state results are schedule independent).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FEC = {"1/2": 0, "2/3": 1, "4/6": 2, "3/4": 3, "5/6": 4, "7/8": 5}
CSTLN = {"BPSK": 0, "QPSK": 1, "8PSK": 2, "16APSK": 3, "32APSK": 4, "64APSKe": 5,
         "16QAM": 6, "64QAM": 7, "256QAM": 8}
SAMPLER = {"nearest": 0, "linear": 1, "rrc": 2}
FMT = {"u8": 0, "s8": 1, "u16": 2, "s16": 3, "f32": 4}
FMT_DTYPE = {"u8": np.uint8, "s8": np.int8, "u16": np.uint16, "s16": np.int16,
             "f32": np.float32}


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("dvbs_oracle.c", "dvbs_tx_oracle.c", "dvbs_hs_oracle.c", "dvbs_oracle.h")]
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, sz, u8p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint8)
        L.orc_sizeof.restype = sz
        L.orc_sizeof.argtypes = [C.c_int]
        L.orc_lowpass.restype = C.c_int
        L.orc_lowpass.argtypes = [C.c_int, C.c_float, vp]
        L.orc_resample_design.restype = C.c_int
        L.orc_resample_design.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float,
                                          C.c_uint, vp, C.c_int, C.POINTER(C.c_int)]
        L.orc_rrc.restype = C.c_int
        L.orc_rrc.argtypes = [C.c_int, C.c_float, C.c_float, vp]
        L.orc_deconv_polys.restype = C.c_int
        L.orc_deconv_polys.argtypes = [C.c_int, vp, vp, C.POINTER(C.c_int)]
        L.orc_cstln_build.argtypes = [vp, C.c_int, C.c_int]
        L.orc_cstln_build2.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_cstln_build2.restype = C.c_int
        L.orc_trig16_build.argtypes = [vp]
        L.orc_rs_tables.argtypes = [vp, vp, vp]
        L.orc_derand_pattern.argtypes = [vp]
        L.orc_cconvert.argtypes = [vp, C.c_int, vp, sz]
        L.orc_scale.argtypes = [vp, C.c_float, vp, sz]
        L.orc_rotator_init.argtypes = [vp, C.c_float]
        L.orc_rotator_run.argtypes = [vp, vp, vp, sz]
        L.orc_fir_init.argtypes = [vp, C.c_uint, vp, C.c_uint]
        L.orc_fir_set_freq.argtypes = [vp, C.c_float]
        L.orc_fir_run.restype = sz
        L.orc_fir_run.argtypes = [vp, vp, sz, vp, C.POINTER(sz)]
        L.orc_fir_shifted.restype = C.POINTER(C.c_float)
        L.orc_fir_shifted.argtypes = [vp]
        L.orc_decimate.restype = sz
        L.orc_decimate.argtypes = [vp, sz, C.c_uint, vp, C.POINTER(sz)]
        L.orc_notch_init.argtypes = [vp, C.c_int]
        L.orc_notch_run.restype = sz
        L.orc_notch_run.argtypes = [vp, vp, sz, vp]
        L.orc_meas_init.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_meas_run.restype = sz
        L.orc_meas_run.argtypes = [vp, vp, sz, C.c_float, vp, sz, C.POINTER(sz)]
        L.orc_notch_get_state.argtypes = [vp, vp, vp, vp, vp]
        L.orc_fft_inplace.argtypes = [C.c_int, vp, C.c_int]
        L.orc_rx_init.argtypes = [vp, vp, vp, C.c_int]
        L.orc_rx_set_omega.argtypes = [vp, C.c_float]
        L.orc_rx_set_freq.argtypes = [vp, C.c_float]
        L.orc_rx_set_rrc.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_rx_config.argtypes = [vp, C.c_float, C.c_int, C.c_ulong]
        L.orc_rx_readahead.restype = C.c_int
        L.orc_rx_readahead.argtypes = [vp]
        L.orc_rx_run.restype = sz
        L.orc_rx_run.argtypes = [vp, vp, sz, vp, C.POINTER(sz), vp, C.POINTER(sz),
                                 vp, C.POINTER(sz)]
        L.orc_rx_get_state.argtypes = [vp, vp]
        L.orc_rx_set_state.argtypes = [vp, vp]
        L.orc_rx_get_limits.argtypes = [vp, vp]
        L.orc_deconv_init.argtypes = [vp, C.c_int]
        L.orc_deconv_next_sync.argtypes = [vp]
        L.orc_deconv_run.restype = sz
        L.orc_deconv_run.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
        L.orc_deconv_run2.restype = sz
        L.orc_deconv_run2.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), C.c_int]
        L.orc_deconv_set.argtypes = [vp, C.c_int, C.c_int]
        L.orc_deconv_set_fastlock.argtypes = [vp, C.c_int]
        L.orc_mpegsync_set_fastlock.argtypes = [vp, C.c_int, C.c_int]
        L.orc_deconv_get.argtypes = [vp] + [C.POINTER(C.c_int)] * 4
        L.orc_mpegsync_init.argtypes = [vp]
        L.orc_mpegsync_run.restype = sz
        L.orc_mpegsync_run.argtypes = [vp, vp, vp, sz, vp, sz, C.POINTER(sz),
                                       vp, C.POINTER(sz), vp, C.POINTER(sz)]
        L.orc_mpegsync_run2.restype = sz
        L.orc_mpegsync_run2.argtypes = [vp, vp, vp, sz, vp, sz, C.POINTER(sz),
                                        vp, C.POINTER(sz), vp, C.POINTER(sz), C.c_int, C.POINTER(C.c_int)]
        L.orc_mpegsync_get.argtypes = [vp, vp]
        L.orc_deinterleave.restype = sz
        L.orc_deinterleave.argtypes = [vp, sz, vp, C.POINTER(sz)]
        L.orc_rs_decode_packet.restype = C.c_int
        L.orc_rs_decode_packet.argtypes = [vp, vp, C.POINTER(C.c_int)]
        L.orc_rs_encode.argtypes = [vp]
        L.orc_hsrx_init.argtypes = [vp]
        L.orc_hsrx_set_omega.argtypes = [vp, C.c_float]
        L.orc_hsrx_set_freq.argtypes = [vp, C.c_float]
        L.orc_hsrx_config.argtypes = [vp, C.c_int, C.c_ulong]
        L.orc_hsrx_run.restype = sz
        L.orc_hsrx_run.argtypes = [vp, vp, sz, vp, C.POINTER(sz), vp, C.POINTER(sz)]
        L.orc_hsdeconv_init.argtypes = [vp, C.c_int]
        L.orc_hsdeconv_run.restype = sz
        L.orc_hsdeconv_run.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
        L.orc_tx_chain.restype = sz
        L.orc_tx_chain.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_char_p,
                                   C.c_int, vp, sz, vp, C.POINTER(sz), vp, C.POINTER(sz)]
        L.orc_tx_taps.restype = C.c_int
        L.orc_tx_taps.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, vp]
        L.orc_tx_amp.restype = C.c_float
        L.orc_tx_amp.argtypes = [C.c_char_p]
        L.orc_tx_resample.restype = sz
        L.orc_tx_resample.argtypes = [vp, sz, vp, C.c_int, C.c_int, vp, C.POINTER(sz)]
        L.orc_tx_randomize.argtypes = [vp, sz, vp]
        L.orc_tx_rs_encode.argtypes = [vp, sz, vp]
        L.orc_derandomize.restype = sz
        L.orc_derandomize.argtypes = [vp, vp, sz, vp]
        L.orc_viterbi_new.restype = vp
        L.orc_viterbi_new.argtypes = [vp, C.c_int]
        L.orc_viterbi_free.argtypes = [vp]
        L.orc_viterbi_set_resync_period.argtypes = [vp, C.c_int]
        L.orc_viterbi_nsyncs.restype = C.c_int
        L.orc_viterbi_nsyncs.argtypes = [vp]
        L.orc_viterbi_current_sync.restype = C.c_int
        L.orc_viterbi_current_sync.argtypes = [vp]
        L.orc_viterbi_run.restype = sz
        L.orc_viterbi_run.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _Obj:
    """Opaque C struct held in a zeroed numpy buffer."""

    def __init__(self, what: int):
        self.buf = np.zeros(lib().orc_sizeof(what) + 64, dtype=np.uint8)

    @property
    def p(self):
        return _p(self.buf)


# ----------------------------------------------------------------- tables

def cstln_table(kind: str = "QPSK", harden: bool = False, fec: str = "3/4"):
    """-> (cells int16[256,256,4] = cost,symbol,phase_error,0 ; symbols int8[n,2]).
    fec only matters for the APSK ring ratios (make_dvbs2_constellation, dvb.h:45-81)."""
    o = _Obj(0)
    if lib().orc_cstln_build2(o.p, CSTLN[kind], FEC[fec], int(harden)):
        raise ValueError(f"Code rate {fec} not supported with {kind}")
    cells = o.buf[:256 * 256 * 8].view(np.int16).reshape(256, 256, 4).copy()
    off = 256 * 256 * 8
    sre = o.buf[off:off + 256].view(np.int8)
    sim = o.buf[off + 256:off + 512].view(np.int8)
    nsym = int(o.buf[off + 512:off + 516].view(np.int32)[0])
    syms = np.stack([sre[:nsym], sim[:nsym]], axis=1).copy()
    return cells, syms, o


def trig16_table() -> np.ndarray:
    t = np.zeros((65536, 2), dtype=np.float32)
    lib().orc_trig16_build(_p(t))
    return t


def rs_tables():
    e = np.zeros(511, np.uint8); l = np.zeros(256, np.uint8); g = np.zeros(17, np.uint8)
    lib().orc_rs_tables(_p(e), _p(l), _p(g))
    return e, l, g


def derand_pattern() -> np.ndarray:
    p = np.zeros(1504, np.uint8)
    lib().orc_derand_pattern(_p(p))
    return p


def lowpass(order: int, fcut: float) -> np.ndarray:
    c = np.zeros(order + 1, np.float32)
    n = lib().orc_lowpass(order, fcut, _p(c))
    return c[:n]


def resample_design(Fs, Fm, rolloff=0.35, rej=10.0, decim=0):
    c = np.zeros(8192, np.float32)
    d = C.c_int(0)
    n = lib().orc_resample_design(Fs, Fm, rolloff, rej, decim, _p(c), c.size, C.byref(d))
    if n < 0:
        raise ValueError("filter too long")
    return c[:n].copy(), d.value


def rrc_design(Fs, Fm, rolloff=0.35, rej=10.0, steps=0):
    """leandvb.cc:437-456 -> (coeffs, steps)"""
    f32 = np.float32
    Fs, Fm, rolloff, rej = f32(Fs), f32(Fm), f32(rolloff), f32(rej)
    if steps == 0:
        steps = max(1, int(f32(64) * Fm / Fs))
    Frrc = f32(Fs * f32(steps))
    transition = f32(f32(Fm / f32(2)) * rolloff)
    order = int(f32(rej * Frrc) / f32(f32(22) * transition))
    c = np.zeros(order + 3, np.float32)
    n = lib().orc_rrc(order, f32(Fm / Frrc), rolloff, _p(c))
    return c[:n].copy(), steps


def deconv_polys(fec: str):
    d = np.zeros(8, np.uint64); d2 = np.zeros(8, np.uint64); w = C.c_int(0)
    pp = lib().orc_deconv_polys(FEC[fec], _p(d), _p(d2), C.byref(w))
    return d[:pp].copy(), d2[:pp].copy(), pp, w.value


# ----------------------------------------------------------------- stages

def cconvert(raw: np.ndarray, fmt: str) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=FMT_DTYPE[fmt]).reshape(-1)
    n = raw.size // 2
    out = np.empty(2 * n, np.float32)
    lib().orc_cconvert(_p(raw), FMT[fmt], _p(out), n)
    return out


def scale(x: np.ndarray, s: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    out = np.empty_like(x)
    lib().orc_scale(_p(x), s, _p(out), x.size // 2)
    return out


class Rotator:
    def __init__(self, freq: float):
        self.o = _Obj(1)
        lib().orc_rotator_init(self.o.p, freq)

    def run(self, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        out = np.empty_like(x)
        lib().orc_rotator_run(self.o.p, _p(x), _p(out), x.size // 2)
        return out


class Fir:
    def __init__(self, coeffs, decim=1):
        self.o = _Obj(2)
        self.coeffs = np.ascontiguousarray(coeffs, np.float32)
        self.decim = decim
        lib().orc_fir_init(self.o.p, self.coeffs.size, _p(self.coeffs), decim)

    def set_freq(self, f):
        lib().orc_fir_set_freq(self.o.p, f)

    def shifted(self):
        return np.ctypeslib.as_array(lib().orc_fir_shifted(self.o.p),
                                     (self.coeffs.size, 2)).copy()

    def run(self, x):
        """-> (out cf32 flat, consumed samples)"""
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        n = x.size // 2
        out = np.empty(2 * (n // self.decim + 1), np.float32)
        cons = C.c_size_t(0)
        k = lib().orc_fir_run(self.o.p, _p(x), n, _p(out), C.byref(cons))
        return out[:2 * k], cons.value


def decimate(x, d):
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    out = np.empty(2 * (x.size // 2 // d + 1), np.float32)
    cons = C.c_size_t(0)
    k = lib().orc_decimate(_p(x), x.size // 2, d, _p(out), C.byref(cons))
    return out[:2 * k], cons.value


class Notch:
    def __init__(self, nslots=1):
        self.o = _Obj(3)
        self.nslots = nslots
        lib().orc_notch_init(self.o.p, nslots)

    def run(self, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        out = np.empty_like(x)
        k = lib().orc_notch_run(self.o.p, _p(x), x.size // 2, _p(out))
        return out[:2 * k], k

    def state(self):
        ph = C.c_int32(0); g = C.c_float(0)
        si = np.zeros(self.nslots, np.int32); es = np.zeros((self.nslots, 2), np.float32)
        lib().orc_notch_get_state(self.o.p, C.byref(ph), C.byref(g), _p(si), _p(es))
        return {"phase": ph.value, "gain": g.value, "slot_i": si, "estim": es}


class Meas:
    """cnr_fft (n = 4096, bandwidth = Fm/Fs) or spectrum (n = 1024, bandwidth = 0)."""

    def __init__(self, n, bandwidth, kavg, decimation):
        self.o = _Obj(8)
        self.n = n
        self.rows = bandwidth == 0
        lib().orc_meas_init(self.o.p, C.c_int(n), C.c_float(bandwidth), C.c_float(kavg), C.c_int(decimation))

    def run(self, x, center_freq=0.0):
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        cap = x.size // 2 // self.n + 1
        out = np.zeros(cap * (self.n if self.rows else 1), np.float32)
        used = C.c_size_t(0)
        k = lib().orc_meas_run(self.o.p, _p(x), C.c_size_t(x.size // 2), C.c_float(center_freq), _p(out),
                               C.c_size_t(cap), C.byref(used))
        return (out[:k * self.n].reshape(k, self.n) if self.rows else out[:k]), used.value


def fft_inplace(x, reverse=True):
    x = np.ascontiguousarray(x, np.float32).reshape(-1).copy()
    lib().orc_fft_inplace(x.size // 2, _p(x), int(reverse))
    return x


class Receiver:
    def __init__(self, cstln_obj, trig, sampler="linear"):
        self.o = _Obj(4)
        self._cst = cstln_obj
        self._trig = trig
        lib().orc_rx_init(self.o.p, cstln_obj.p, _p(trig), SAMPLER[sampler])

    def set_omega(self, w): lib().orc_rx_set_omega(self.o.p, w)
    def set_freq(self, f): lib().orc_rx_set_freq(self.o.p, f)

    def set_rrc(self, coeffs, sub):
        self._rrc = np.ascontiguousarray(coeffs, np.float32)
        lib().orc_rx_set_rrc(self.o.p, self._rrc.size, _p(self._rrc), sub)

    def config(self, pll_adjustment=1.0, allow_drift=False, meas_decimation=1048576):
        lib().orc_rx_config(self.o.p, pll_adjustment, int(allow_drift), meas_decimation)

    def readahead(self): return lib().orc_rx_readahead(self.o.p)

    def get_state(self):
        w = np.zeros(22, np.uint32); lib().orc_rx_get_state(self.o.p, _p(w)); return w

    def set_state(self, w):
        w = np.ascontiguousarray(w, np.uint32); lib().orc_rx_set_state(self.o.p, _p(w))

    def limits(self):
        f = np.zeros(4, np.float32); lib().orc_rx_get_limits(self.o.p, _p(f)); return f

    def run(self, x):
        """-> dict(symbols uint8[n,4], sampled cf32, meas f32[m,3], consumed)"""
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        n = x.size // 2
        sym = np.zeros((n + 256, 4), np.uint8)
        smp = np.zeros((n // 128 + 2, 2), np.float32)
        meas = np.zeros((n // 128 + 2, 3), np.float32)
        ns, nsm, nm = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        cons = lib().orc_rx_run(self.o.p, _p(x), n, _p(sym), C.byref(ns), _p(smp),
                                C.byref(nsm), _p(meas), C.byref(nm))
        return {"symbols": sym[:ns.value], "sampled": smp[:nsm.value],
                "meas": meas[:nm.value], "consumed": cons}


class Deconv:
    def __init__(self, fec="1/2", fastlock=False):
        self.o = _Obj(5)
        lib().orc_deconv_init(self.o.p, FEC[fec])
        self.fastlock = bool(fastlock)
        if fastlock:
            lib().orc_deconv_set_fastlock(self.o.p, 1)

    def next_sync(self): lib().orc_deconv_next_sync(self.o.p)

    def get(self):
        v = [C.c_int(0) for _ in range(4)]
        lib().orc_deconv_get(self.o.p, *[C.byref(x) for x in v])
        return {"locked": v[0].value, "skip": v[1].value,
                "punctperiod": v[2].value, "punctweight": v[3].value}

    def run(self, symbols4, out_cap=None, big_batch=False):
        """One reference run(); big_batch drops the `n < 32` early return (dvb.h:424-426),
        which only matters for the last few bytes of a stream."""
        s = np.ascontiguousarray(symbols4, np.uint8).reshape(-1, 4)
        cap = out_cap if out_cap is not None else s.shape[0] + 64
        out = np.zeros(max(cap, 1), np.uint8)
        cons = C.c_size_t(0)
        k = lib().orc_deconv_run2(self.o.p, _p(s), s.shape[0], _p(out), cap, C.byref(cons), int(big_batch))
        return out[:k], cons.value

    def set_locked(self, locked, skip):
        lib().orc_deconv_set(self.o.p, locked, skip)

    def snapshot(self): return self.o.buf.copy()
    def restore(self, b): self.o.buf[:] = b


class MpegSync:
    def __init__(self, fastlock=False, resync_period=1):
        self.o = _Obj(6)
        lib().orc_mpegsync_init(self.o.p)
        if fastlock:
            lib().orc_mpegsync_set_fastlock(self.o.p, 1, resync_period)

    def get(self):
        v = np.zeros(7, np.int64); lib().orc_mpegsync_get(self.o.p, _p(v))
        return dict(zip(["bitphase", "polarity", "synchronized", "phase8",
                         "next_sync_count", "lock_timeleft", "locktime"], v.tolist()))

    def run(self, data, deconv=None, out_cap=None, per_wrap=False):
        """-> (aligned bytes, consumed, lock events, locktimes[, switched when per_wrap])"""
        d = np.ascontiguousarray(data, np.uint8)
        cap = out_cap if out_cap is not None else d.size + 204 * 9
        out = np.zeros(cap, np.uint8)
        lock = np.zeros(d.size // 204 + 8, np.int32)
        lt = np.zeros(d.size // 204 + 8, np.uint64)
        cons, nl, nlt = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        sw = C.c_int(0)
        k = lib().orc_mpegsync_run2(self.o.p, deconv.o.p if deconv else None, _p(d), d.size,
                                    _p(out), cap, C.byref(cons), _p(lock), C.byref(nl),
                                    _p(lt), C.byref(nlt), int(per_wrap), C.byref(sw))
        res = (out[:k], cons.value, lock[:nl.value].copy(), lt[:nlt.value].copy())
        return res + (bool(sw.value),) if per_wrap else res


def deinterleave(data):
    d = np.ascontiguousarray(data, np.uint8)
    out = np.zeros((d.size // 204 + 1, 204), np.uint8)
    cons = C.c_size_t(0)
    k = lib().orc_deinterleave(_p(d), d.size, _p(out), C.byref(cons))
    return out[:k], cons.value


def rs_decode(packets204):
    """-> (ts188 [n,188], corrupted bool[n], bits_corrected int[n], fixed204)"""
    p = np.ascontiguousarray(packets204, np.uint8).reshape(-1, 204).copy()
    n = p.shape[0]
    out = np.zeros((n, 188), np.uint8)
    bad = np.zeros(n, bool); nerr = np.zeros(n, np.int32)
    for k in range(n):
        e = C.c_int(0)
        bad[k] = bool(lib().orc_rs_decode_packet(_p(p[k]), _p(out[k]), C.byref(e)))
        nerr[k] = e.value
    return out, bad, nerr, p


def rs_encode(msgs188):
    m = np.ascontiguousarray(msgs188, np.uint8).reshape(-1, 188)
    out = np.zeros((m.shape[0], 204), np.uint8)
    out[:, :188] = m
    for k in range(m.shape[0]):
        lib().orc_rs_encode(_p(out[k]))
    return out


class Derand:
    def __init__(self):
        self.o = _Obj(7)

    def run(self, packets188):
        p = np.ascontiguousarray(packets188, np.uint8).reshape(-1, 188)
        out = np.zeros_like(p)
        k = lib().orc_derandomize(self.o.p, _p(p), p.shape[0], _p(out))
        return out[:k]


class Viterbi:
    def __init__(self, cstln_obj, fec="1/2"):
        self._cst = cstln_obj
        self.h = lib().orc_viterbi_new(cstln_obj.p, FEC[fec])
        if not self.h:
            raise ValueError("unsupported code rate")

    def __del__(self):
        try:
            lib().orc_viterbi_free(self.h)
        except Exception:
            pass

    def set_resync_period(self, p): lib().orc_viterbi_set_resync_period(self.h, p)
    def nsyncs(self): return lib().orc_viterbi_nsyncs(self.h)
    def current_sync(self): return lib().orc_viterbi_current_sync(self.h)

    def run(self, symbols4):
        s = np.ascontiguousarray(symbols4, np.uint8).reshape(-1, 4)
        out = np.zeros(s.shape[0] + 64, np.uint8)
        cons = C.c_size_t(0)
        k = lib().orc_viterbi_run(self.h, _p(s), s.shape[0], _p(out), out.size, C.byref(cons))
        return out[:k], cons.value


# ------------------------------------------------------------------ chain

@dataclass
class Config:
    """Mirror of the leandvb flags that reach the hot path (leandvb.cc:43-136)."""
    fmt: str = "u8"
    float_scale: float = 1.0
    Fs: float = 2.4e6
    Fm: float = 2e6
    anf: int = 1
    Fderot: float = 0.0
    resample: bool = False
    resample_rej: float = 10.0
    decim: int = 0
    sampler: str = "linear"
    rrc_steps: int = 0
    rrc_rej: float = 10.0
    rolloff: float = 0.35
    cstln: str = "QPSK"
    fec: str = "1/2"
    viterbi: bool = False
    hard_metric: bool = False
    fastlock: bool = False
    allow_drift: bool = False
    Ftune: float = 0.0
    Finfo: float = 5.0
    cnr: bool = False          # --cnr (leandvb.cc:322-329); the spectrum is always measured (:333-343)


def _idecim(a, b):
    d = int(np.float32(a) / np.float32(b))
    return max(d, 1)


class Chain:
    """One-shot stage-by-stage run of the whole receive path on the CPU."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        f32 = np.float32
        self.cells, self.syms, self.cst = cstln_table(cfg.cstln, cfg.hard_metric, cfg.fec)
        self.trig = trig16_table()
        self.notch = Notch(cfg.anf) if cfg.anf else None
        Fs = f32(cfg.Fs)
        self.rot = Rotator(f32(-f32(cfg.Fderot) / Fs)) if cfg.Fderot else None
        self.fir = None
        self.decim = 1
        if cfg.resample:
            taps, d = resample_design(cfg.Fs, cfg.Fm, cfg.rolloff, cfg.resample_rej, cfg.decim)
            self.fir_taps = taps
            self.fir = Fir(taps, d)
            self.decim = d
            Fs = f32(Fs / f32(d))
        elif cfg.decim > 1:
            self.decim = cfg.decim
            Fs = f32(Fs / f32(cfg.decim))
        self.Fs_rx = Fs
        self.rx = Receiver(self.cst, self.trig, cfg.sampler)
        if cfg.sampler == "rrc":
            c, steps = rrc_design(Fs, cfg.Fm, cfg.rolloff, cfg.rrc_rej, cfg.rrc_steps)
            self.rrc_taps, self.rrc_steps = c, steps
            self.rx.set_rrc(c, steps)
        self.rx.set_omega(f32(Fs / f32(cfg.Fm)))
        if cfg.Ftune:
            self.rx.set_freq(f32(f32(cfg.Ftune) / Fs))
        pll = f32(1.0)
        if cfg.viterbi:
            pll = f32(pll / f32(6))
        self.rx.config(pll, cfg.allow_drift, _idecim(Fs, cfg.Finfo))
        fec = cfg.fec
        if cfg.viterbi and fec == "2/3" and self.syms.shape[0] in (4, 64):   # leandvb.cc:533-537
            fec = "4/6"
        self.vit = Viterbi(self.cst, fec) if cfg.viterbi else None
        if self.vit and cfg.fastlock:
            self.vit.set_resync_period(1)                      # leandvb.cc:540
        self.deconv = None if cfg.viterbi else Deconv(fec, fastlock=cfg.fastlock)
        self.sync = MpegSync(fastlock=cfg.fastlock, resync_period=1)   # leandvb.cc:553, 565
        self.derand = Derand()

    def run(self, raw: np.ndarray) -> dict:
        cfg = self.cfg
        t = {}
        if cfg.fmt == "f32":
            x = scale(np.ascontiguousarray(raw, np.float32), np.float32(cfg.float_scale))
        else:
            x = cconvert(raw, cfg.fmt)
        t["rawiq"] = x
        if self.notch:
            x, _ = self.notch.run(x)
            t["notched"] = x
        if self.rot:
            x = self.rot.run(x)
        # cnr_fft / spectrum read the stream in front of the FIR (leandvb.cc:322-343);
        # freq_tap is the demodulator's value when the batch starts (0 after a reset)
        dec1 = _idecim(self.cfg.Fs, 1)
        if self.cfg.cnr:
            bw = np.float32(self.cfg.Fm) / np.float32(self.cfg.Fs)
            t["cnr"], _ = Meas(4096, float(bw), 0.1, dec1).run(x, 0.0)
        t["spectrum"], _ = Meas(1024, 0.0, 0.5, dec1).run(x)
        if self.fir:
            # fir_filter::run (dsp.h:236-244) follows the demodulator's freq_tap when it is further than
            # freq_tol from the frequency the taps are shifted for (leandvb.cc:505-510).  One-shot schedule:
            # sampled once, before the batch (the reference samples it at every run() call) -- with --tune
            # the very first call already retunes.
            f32 = np.float32
            freq_tap = f32(self.rx.get_state().view(f32)[20])
            new_freq = f32(freq_tap * f32(1.0 / self.decim))
            tol = f32(np.float64(f32(cfg.Fm) / (f32(self.Fs_rx) * f32(self.decim))) * 0.1)
            if abs(np.float64(f32(0.0) - new_freq)) > tol:
                self.fir.set_freq(float(new_freq))
            x, _ = self.fir.run(x)
        elif self.decim > 1:
            x, _ = decimate(x, self.decim)
        t["pp"] = x
        r = self.rx.run(x)
        t["symbols"] = r["symbols"]; t["sampled"] = r["sampled"]; t["meas"] = r["meas"]
        if self.vit:
            by, _ = self.vit.run(r["symbols"])
            t["bytes"] = by
            mp, lock, lt = self._sync_all(by, None)
        elif cfg.fastlock:
            by, mp, lock, lt = self._deconv_sync_fastlock(r["symbols"])
            t["bytes"] = by
        else:
            by, mp, lock, lt = self._deconv_sync(r["symbols"])
            t["bytes"] = by
        t["mpegbytes"] = mp; t["lock"] = lock; t["locktime"] = lt
        rsp, _ = deinterleave(mp)
        t["rspackets"] = rsp
        ts, bad, nerr, _ = rs_decode(rsp)
        t["rtspackets"] = ts; t["rs_bad"] = bad; t["rs_nerr"] = nerr
        t["ts"] = self.derand.run(ts)
        return t

    def _sync_all(self, by, deconv):
        outs, locks, lts = [], [], []
        pos = 0
        while True:
            o, c, lk, lt = self.sync.run(by[pos:], deconv)
            outs.append(o); locks.append(lk); lts.append(lt)
            if c == 0 and o.size == 0:
                break
            pos += c
        return np.concatenate(outs), np.concatenate(locks), np.concatenate(lts)

    FASTLOCK_PROBE = 1024

    def _deconv_sync_fastlock(self, symbols):
        """--fastlock under the large-batch schedule: one deconvol_sync::run() per
        window (dvb.h:414-467), a window being everything the batch allows -- cut to FASTLOCK_PROBE
        bytes while even the best alignment is wrong on more than a third of the bits, so that the
        one-symbol skip (dvb.h:449-453) acts after a window of the size the reference sees with
        its default buffers; mpeg_sync searches with run_searching_fast (dvb.h:781-796)."""
        by_all, outs, locks, lts = [], [], [], []
        spos = 0
        bbuf = np.zeros(0, np.uint8)
        for _ in range(1 << 20):
            snap = self.deconv.snapshot()
            by, cons = self.deconv.run(symbols[spos:], big_batch=True)
            if by.size == 0 and cons == 0:
                break
            if self.deconv.get()["skip"] and by.size > self.FASTLOCK_PROBE:
                self.deconv.restore(snap)
                by, cons = self.deconv.run(symbols[spos:], out_cap=self.FASTLOCK_PROBE, big_batch=True)
            spos += cons
            by_all.append(by)
            bbuf = np.concatenate([bbuf, by])
            while True:
                o, c, lk, lt = self.sync.run(bbuf, None)
                outs.append(o); locks.append(lk); lts.append(lt)
                bbuf = bbuf[c:]
                if c == 0 and o.size == 0:
                    break
        by = np.concatenate(by_all) if by_all else np.zeros(0, np.uint8)
        return (by, np.concatenate(outs) if outs else np.zeros(0, np.uint8),
                np.concatenate(locks) if locks else np.zeros(0, np.int32),
                np.concatenate(lts) if lts else np.zeros(0, np.uint64))

    def _deconv_sync(self, symbols):
        """Algebraic deconvolution + MPEG sync with the backward next_sync() edge
        (dvb.h:771-778) under the large-batch schedule: the sweep counter
        advances per wrap of the bit phase; when it fires, the bytes that the old
        hypothesis produced beyond the search position are void and the next
        hypothesis starts at the symbol that follows the last consumed byte."""
        by_all, outs, locks, lts = [], [], [], []
        spos = 0
        for _ in range(256):
            if self.deconv.get()["skip"]:
                if len(symbols) - spos < 1:
                    break
            snap = self.deconv.snapshot()
            by, scons = self.deconv.run(symbols[spos:], big_batch=True)
            if by.size == 0 and not self.deconv.get()["skip"]:
                break
            bpos = 0
            switched = False
            while True:
                o, c, lk, lt, sw = self.sync.run(by[bpos:], self.deconv, per_wrap=True)
                outs.append(o); locks.append(lk); lts.append(lt)
                bpos += c
                if sw:
                    switched = True
                    break
                if c == 0 and o.size == 0:
                    break
            if not switched:
                by_all.append(by)
                break
            st = self.deconv.get()
            self.deconv.restore(snap)              # back to the state before this run
            used = bpos
            by2, scons2 = self.deconv.run(symbols[spos:], out_cap=used, big_batch=True) if used else (by[:0], 0)
            by_all.append(by2)
            spos += scons2
            self.deconv.set_locked(st["locked"], st["skip"])
        by = np.concatenate(by_all) if by_all else np.zeros(0, np.uint8)
        return (by, np.concatenate(outs) if outs else np.zeros(0, np.uint8),
                np.concatenate(locks) if locks else np.zeros(0, np.int32),
                np.concatenate(lts) if lts else np.zeros(0, np.uint64))


# ------------------------------------------------------------------ --hs path

class HsReceiver:
    """fast_qpsk_receiver<u8> (sdr.h:946-1189)."""
    def __init__(self, omega, freq=0.0, allow_drift=False, meas_decimation=1048576):
        self.o = _Obj(9)
        lib().orc_hsrx_init(self.o.p)
        lib().orc_hsrx_set_omega(self.o.p, np.float32(omega))
        if freq:
            lib().orc_hsrx_set_freq(self.o.p, np.float32(freq))
        lib().orc_hsrx_config(self.o.p, int(allow_drift), int(meas_decimation))

    def run(self, raw_u8):
        x = np.ascontiguousarray(raw_u8, np.uint8).reshape(-1)
        n = x.size // 2
        sym = np.zeros(n + 256, np.uint8)
        freq = np.zeros(n // 128 + 8, np.float32)
        ns, nf = C.c_size_t(0), C.c_size_t(0)
        used = lib().orc_hsrx_run(self.o.p, _p(x), n, _p(sym), C.byref(ns), _p(freq), C.byref(nf))
        return sym[: ns.value].copy(), freq[: nf.value].copy(), used


class HsDeconv:
    """dvb_deconvol_sync_hard (dvb.h:612-707)."""
    def __init__(self, resync_period=32):
        self.o = _Obj(10)
        lib().orc_hsdeconv_init(self.o.p, resync_period)

    def run(self, sym):
        s = np.ascontiguousarray(sym, np.uint8)
        out = np.zeros(s.size // 8 + 64, np.uint8)
        cons = C.c_size_t(0)
        k = lib().orc_hsdeconv_run(self.o.p, _p(s), s.size, _p(out), out.size, C.byref(cons))
        return out[:k].copy(), cons.value


def hs_chain(raw_u8, Fs=2.4e6, Fm=2e6, fastlock=False, Ftune=0.0, allow_drift=False, Finfo=5.0) -> dict:
    """run_highspeed (apps/leandvb.cc:727-969) stage by stage on a whole u8 IQ array."""
    f32 = np.float32
    period = 1 if fastlock else 32                                  # leandvb.cc:853, 863
    rx = HsReceiver(f32(f32(Fs) / f32(Fm)), f32(f32(Ftune) / f32(Fs)) if Ftune else 0.0, allow_drift,
                    _idecim(f32(Fs), Finfo))
    t = {}
    t["symbols"], t["freq"], _ = rx.run(raw_u8)
    t["bytes"], _ = HsDeconv(period).run(t["symbols"])
    sync = MpegSync(fastlock=True, resync_period=period)
    outs, bbuf = [], t["bytes"]
    while True:
        o, c, _, _ = sync.run(bbuf, None)
        outs.append(o); bbuf = bbuf[c:]
        if c == 0 and o.size == 0:
            break
    t["mpegbytes"] = np.concatenate(outs)
    t["rspackets"], _ = deinterleave(t["mpegbytes"])
    ts, bad, nerr, _ = rs_decode(t["rspackets"])
    t["rtspackets"] = ts; t["rs_bad"] = bad; t["rs_nerr"] = nerr
    t["ts"] = Derand().run(ts)
    return t


# ------------------------------------------------------------------ transmit chain

def tx_taps(interp: int, rolloff: float = 0.35, rrc_rej: float = 10.0, power: str = "0") -> np.ndarray:
    """leandvbtx.cc:131-138: RRC interpolation taps after normalize_power."""
    L = lib()
    out = np.zeros(int(interp * rrc_rej) + 8, np.float32)
    n = L.orc_tx_taps(interp, rolloff, rrc_rej, L.orc_tx_amp(str(power).encode()), _p(out))
    return out[:n].copy()


def tx_resample(x: np.ndarray, coeffs: np.ndarray, interp: int) -> np.ndarray:
    """fir_resampler<cf32,float> with real coefficients at frequency 0 (dsp.h:290-364)."""
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    c = np.ascontiguousarray(coeffs, np.float32)
    y = np.zeros(2 * (x.size // 2 * interp + 8), np.float32)
    used = C.c_size_t(0)
    n = lib().orc_tx_resample(_p(x), x.size // 2, _p(c), c.size, interp, _p(y), C.byref(used))
    return y[: 2 * n].copy()


def tx_rs_packets(ts: np.ndarray) -> np.ndarray:
    """randomizer + rs_encoder (dvb.h:1063-1102, 957-980): [n,188] -> [n,204]."""
    ts = np.ascontiguousarray(ts, np.uint8).reshape(-1, 188)
    r = np.zeros_like(ts)
    lib().orc_tx_randomize(_p(ts), ts.shape[0], _p(r))
    out = np.zeros((ts.shape[0], 204), np.uint8)
    lib().orc_tx_rs_encode(_p(r), ts.shape[0], _p(out))
    return out


def tx_chain(ts: np.ndarray, cstln: str = "QPSK", cr: str = "1/2", ratio: str = "2", power: str = "0",
             agc: bool = False, rolloff: float = 0.35, rrc_rej: float = 10.0) -> dict:
    """The whole leandvbtx chain (apps/leandvbtx.cc:79-197) on TS packets [n,188]:
    {"iq": interleaved floats, "mpegbytes": u8, "symbols": u8}."""
    ts = np.ascontiguousarray(ts, np.uint8).reshape(-1, 188)
    npk = ts.shape[0]
    parts = str(ratio).split("/")
    I, D = int(parts[0]), (int(parts[1]) if len(parts) > 1 else 1)
    cap = npk * 204 * 16 * I // D + 4096
    iq = np.zeros(2 * cap, np.float32)
    mb = np.zeros(npk * 204 + 16, np.uint8)
    sym = np.zeros(npk * 204 * 16 + 64, np.uint8)
    nmb, nsym = C.c_size_t(0), C.c_size_t(0)
    n = lib().orc_tx_chain(_p(ts), npk, CSTLN[cstln], FEC[cr], I, D, rolloff, rrc_rej, str(power).encode(), int(agc),
                           _p(iq), cap, _p(mb), C.byref(nmb), _p(sym), C.byref(nsym))
    return {"iq": iq[: 2 * n].copy(), "mpegbytes": mb[: nmb.value].copy(), "symbols": sym[: nsym.value].copy()}


def ref_bin(name: str) -> str:
    """Path of a reference binary built by `make -C oracle ref`."""
    return os.path.join(_HERE, "_ref", name)
