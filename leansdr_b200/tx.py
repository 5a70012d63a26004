"""ctypes binding of the transmit chain (include/leandvb_b200_tx.h): the B200-native
leandvbtx.  Like capi.py this only loads libleandvb_b200.so and mirrors the structs; there
is no fallback path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import ABI_VERSION, CSTLN, FEC, LdvbError, _p, load

TX_EXPORTS = [
    "ldvbtx_config_default", "ldvbtx_create", "ldvbtx_destroy", "ldvbtx_reset", "ldvbtx_last_error",
    "ldvbtx_set_stream", "ldvbtx_max_samples", "ldvbtx_push", "ldvbtx_process_device", "ldvbtx_tsgen_device",
    "ldvbtx_tap", "ldvbtx_taps", "ldvbtx_host_taps", "ldvbtx_fir_resampler_cf32",
]
TX_TAP = {"rspackets": 0, "mpegbytes": 1, "symbols": 2}


class TxConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("constellation", C.c_int32), ("fec", C.c_int32),
        ("interp", C.c_int32), ("decim", C.c_int32), ("rolloff", C.c_float), ("rrc_rej", C.c_float),
        ("power_db", C.c_char * 32), ("agc", C.c_int32), ("device", C.c_int32), ("max_packets", C.c_uint64),
        ("keep_taps", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


_bound = False


def _lib():
    global _bound
    L = load()
    if not _bound:
        vp, sz = C.c_void_p, C.c_size_t
        L.ldvbtx_config_default.argtypes = [C.POINTER(TxConfig)]
        L.ldvbtx_create.argtypes = [C.POINTER(TxConfig), C.POINTER(vp)]
        L.ldvbtx_destroy.argtypes = [vp]
        L.ldvbtx_reset.argtypes = [vp]
        L.ldvbtx_last_error.restype = C.c_char_p
        L.ldvbtx_last_error.argtypes = [vp]
        L.ldvbtx_set_stream.argtypes = [vp, vp]
        L.ldvbtx_max_samples.restype = sz
        L.ldvbtx_max_samples.argtypes = [vp, sz]
        L.ldvbtx_push.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
        L.ldvbtx_process_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
        L.ldvbtx_tsgen_device.argtypes = [vp, C.c_uint64, sz, vp]
        L.ldvbtx_tap.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
        L.ldvbtx_taps.argtypes = [vp, vp, sz, C.POINTER(sz)]
        L.ldvbtx_host_taps.argtypes = [C.POINTER(TxConfig), vp, sz, C.POINTER(sz)]
        L.ldvbtx_fir_resampler_cf32.argtypes = [C.c_int, vp, sz, vp, C.c_uint32, C.c_uint32, vp, sz, C.POINTER(sz)]
        _bound = True
    return L


def tx_config(cstln="QPSK", fec="1/2", ratio="2", power="0", agc=False, rolloff=0.35, rrc_rej=10.0,
              max_packets=4096, device=0, keep_taps=False) -> TxConfig:
    """leandvbtx's command line as a config: --const, --cr, -f INTERP[/DECIM], --power, --agc,
    --roll-off, --rrc-rej (apps/leandvbtx.cc:258-297)."""
    cfg = TxConfig()
    _lib().ldvbtx_config_default(C.byref(cfg))
    assert cfg.abi_version == ABI_VERSION
    parts = str(ratio).split("/")
    cfg.constellation = CSTLN[cstln]
    cfg.fec = FEC[fec]
    cfg.interp = int(parts[0])
    cfg.decim = int(parts[1]) if len(parts) > 1 else 1
    cfg.power_db = str(power).encode()
    cfg.agc = int(bool(agc))
    cfg.rolloff = rolloff
    cfg.rrc_rej = rrc_rej
    cfg.max_packets = max_packets
    cfg.device = device
    cfg.keep_taps = int(bool(keep_taps))
    return cfg


def host_taps(cfg: TxConfig) -> np.ndarray:
    """Interpolation taps from the library's host code; needs no GPU."""
    L = _lib()
    n = C.c_size_t(0)
    rc = L.ldvbtx_host_taps(C.byref(cfg), None, 0, C.byref(n))
    if rc:
        raise LdvbError(rc, "ldvbtx_host_taps")
    out = np.empty(n.value, np.float32)
    rc = L.ldvbtx_host_taps(C.byref(cfg), _p(out), out.size, C.byref(n))
    if rc:
        raise LdvbError(rc, "ldvbtx_host_taps")
    return out


class Transmitter:
    """Host-side mirror of the leandvbtx runnable chain replaced by one handle."""

    def __init__(self, cfg: TxConfig | None = None, **kw):
        self.L = _lib()
        self.cfg = cfg if cfg is not None else tx_config(**kw)
        self.h = C.c_void_p()
        rc = self.L.ldvbtx_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise LdvbError(rc, "ldvbtx_create", self.L.ldvb_strerror(rc).decode())

    def close(self):
        if self.h:
            self.L.ldvbtx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, where):
        if rc != 0:
            raise LdvbError(rc, where, self.L.ldvbtx_last_error(self.h).decode() or self.L.ldvb_strerror(rc).decode())

    def reset(self):
        self._ck(self.L.ldvbtx_reset(self.h), "ldvbtx_reset")

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.ldvbtx_set_stream(self.h, C.c_void_p(cuda_stream)), "ldvbtx_set_stream")

    def max_samples(self, n_packets: int) -> int:
        return self.L.ldvbtx_max_samples(self.h, n_packets)

    def push(self, ts: np.ndarray) -> np.ndarray:
        """TS packets [n,188] (host) -> interleaved I/Q floats (host)."""
        ts = np.ascontiguousarray(ts, dtype=np.uint8).reshape(-1, 188)
        cap = self.max_samples(ts.shape[0])
        out = np.empty(2 * cap, np.float32)
        n = C.c_size_t(0)
        self._ck(self.L.ldvbtx_push(self.h, _p(ts), ts.shape[0], _p(out), cap, C.byref(n)), "ldvbtx_push")
        return out[: 2 * n.value]

    def process_device(self, ts_ptr: int, n_packets: int, iq_ptr: int, cap_samples: int) -> int:
        n = C.c_size_t(0)
        self._ck(self.L.ldvbtx_process_device(self.h, C.c_void_p(ts_ptr), n_packets, C.c_void_p(iq_ptr), cap_samples,
                                              C.byref(n)), "ldvbtx_process_device")
        return n.value

    def tsgen_device(self, first: int, n_packets: int, ts_ptr: int):
        self._ck(self.L.ldvbtx_tsgen_device(self.h, first, n_packets, C.c_void_p(ts_ptr)), "ldvbtx_tsgen_device")

    def tap(self, name: str) -> np.ndarray:
        n = C.c_size_t(0)
        self._ck(self.L.ldvbtx_tap(self.h, TX_TAP[name], None, 0, C.byref(n)), "ldvbtx_tap")
        out = np.empty(n.value, np.uint8)
        self._ck(self.L.ldvbtx_tap(self.h, TX_TAP[name], _p(out), out.size, C.byref(n)), "ldvbtx_tap")
        return out

    def taps(self) -> np.ndarray:
        n = C.c_size_t(0)
        self._ck(self.L.ldvbtx_taps(self.h, None, 0, C.byref(n)), "ldvbtx_taps")
        out = np.empty(n.value, np.float32)
        self._ck(self.L.ldvbtx_taps(self.h, _p(out), out.size, C.byref(n)), "ldvbtx_taps")
        return out


def fir_resampler_cf32(x: np.ndarray, taps_cplx: np.ndarray, interp: int, device: int = 0) -> np.ndarray:
    """Stand-alone fir_resampler<cf32,float> (dsp.h:290-364) on the GPU: host in, host out."""
    L = _lib()
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    t = np.ascontiguousarray(taps_cplx, np.float32).reshape(-1)
    n_in, ntaps = x.size // 2, t.size // 2
    cap = n_in * interp + 16
    y = np.empty(2 * cap, np.float32)
    n = C.c_size_t(0)
    rc = L.ldvbtx_fir_resampler_cf32(device, _p(x), n_in, _p(t), ntaps, interp, _p(y), cap, C.byref(n))
    if rc:
        raise LdvbError(rc, "ldvbtx_fir_resampler_cf32", L.ldvb_strerror(rc).decode())
    return y[: 2 * n.value]
