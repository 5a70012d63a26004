"""ctypes binding of libleandvb_b200.so (include/leandvb_b200.h).

The library is the product: hand-written sm_100a kernels behind a C ABI.  This
module only loads it and mirrors the structs; there is NO Python/NumPy/torch
fallback -- if the shared object is missing or no B200 is present the calls
fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDVB_LIB") or os.path.join(_HERE, "libleandvb_b200.so")   # LDVB_LIB: experiment builds only

ABI_VERSION = 3
FMT = {"u8": 0, "s8": 1, "u16": 2, "s16": 3, "f32": 4}
FMT_DTYPE = {"u8": np.uint8, "s8": np.int8, "u16": np.uint16, "s16": np.int16, "f32": np.float32}
CSTLN = {"BPSK": 0, "QPSK": 1, "8PSK": 2, "16APSK": 3, "32APSK": 4, "64APSKe": 5,
         "16QAM": 6, "64QAM": 7, "256QAM": 8}
FEC = {"1/2": 0, "2/3": 1, "4/6": 2, "3/4": 3, "5/6": 4, "7/8": 5}
SAMPLER = {"nearest": 0, "linear": 1, "rrc": 2}
RX_EXACT, RX_FAST = 0, 1
TAP = {"pp": 0, "symbols": 1, "bytes": 2, "mpegbytes": 3, "rspackets": 4, "rtspackets": 5,
       "rsflags": 6, "sampled": 7, "meas": 8}
TABLE = {"cstln": 0, "trig16": 1, "rs_exp": 2, "rs_log": 3, "derand": 4, "fir": 5, "rrc": 6,
         "deconv": 7, "trellis": 8, "vitmap": 9, "hs_polar": 10, "hs_rect": 11, "hs_sincos": 12,
         "fir_shifted": 13}


class LdvbError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: error {code}: {detail}")


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("input_format", C.c_int32), ("float_scale", C.c_float),
        ("Fs", C.c_float), ("Fm", C.c_float), ("anf", C.c_int32), ("Fderot", C.c_float),
        ("resample", C.c_int32), ("resample_rej", C.c_float), ("decim", C.c_uint32),
        ("sampler", C.c_int32), ("rrc_steps", C.c_int32), ("rrc_rej", C.c_float),
        ("rolloff", C.c_float), ("constellation", C.c_int32), ("fec", C.c_int32),
        ("viterbi", C.c_int32), ("hard_metric", C.c_int32), ("fastlock", C.c_int32),
        ("allow_drift", C.c_int32), ("Ftune", C.c_float), ("Finfo", C.c_float),
        ("rx_mode", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_uint64),
        ("span_chunks", C.c_uint32), ("warmup_chunks", C.c_uint32), ("keep_taps", C.c_int32),
        ("push_sub_batch", C.c_int32), ("cnr", C.c_int32), ("spectrum", C.c_int32), ("vber", C.c_int32), ("hs", C.c_int32), ("vit_segments", C.c_int32), ("vit_warm_chunks", C.c_int32),
        ("settle_chunks", C.c_int32), ("seam_mode", C.c_int32), ("async_push", C.c_int32), ("reserved0", C.c_int32),
    ]


class Meas(C.Structure):
    _fields_ = [
        ("freq_tap", C.c_float), ("ss", C.c_float), ("mer", C.c_float), ("lock", C.c_int32),
        ("locktime", C.c_uint64), ("rs_bits", C.c_uint64), ("rs_errs", C.c_uint64),
        ("ts_packets", C.c_uint64), ("ts_dropped", C.c_uint64), ("samples_in", C.c_uint64),
        ("symbols", C.c_uint64), ("seams_total", C.c_uint32), ("seams_repaired", C.c_uint32),
        ("notch_repaired", C.c_uint32), ("kernel_launches", C.c_uint32),
        ("vit_segments", C.c_uint32), ("vit_repaired", C.c_uint32),
        ("seams_mismatch_accepted", C.c_uint32), ("settle_passes", C.c_uint32),
        ("seam_max_dphase", C.c_float), ("seam_max_dfreqw", C.c_float), ("seam_max_dmu", C.c_float),
    ]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


EXPORTS = [
    "ldvb_abi_version", "ldvb_strerror", "ldvb_last_error", "ldvb_config_default", "ldvb_create",
    "ldvb_destroy", "ldvb_push", "ldvb_pull", "ldvb_flush", "ldvb_host_register", "ldvb_host_unregister", "ldvb_process_device", "ldvb_get_meas", "ldvb_tap",
    "ldvb_table", "ldvb_host_table", "ldvb_state_size", "ldvb_get_state", "ldvb_set_state", "ldvb_get_rx_state",
    "ldvb_set_rx_state", "ldvb_fir_cf32", "ldvb_deint_rs", "ldvb_rs_decode",
    "ldvb_reset", "ldvb_set_stream", "ldvb_profile", "ldvb_get_profile",
    "ldvb_pull_cnr", "ldvb_pull_spectrum", "ldvb_pull_vber",
    "ldvb_edge_size", "ldvb_shard_min_halo", "ldvb_shard_detect", "ldvb_shard_front", "ldvb_shard_back",
    "ldvb_ring_unique_id", "ldvb_ring_init", "ldvb_ring_round", "ldvb_ring_flush", "ldvb_ring_stats", "ldvb_ring_destroy",
]


class Shard(C.Structure):
    """ldvb_shard (include/leandvb_b200.h): one time chunk [halo | chunk] of the stream."""
    _fields_ = [("iq_dev", C.c_void_p), ("abs_raw0", C.c_uint64), ("n_halo", C.c_uint64), ("n_chunk", C.c_uint64),
                ("n_halo_next", C.c_uint64), ("last", C.c_int32), ("reserved", C.c_int32),
                ("bins_before", C.c_int32 * 4), ("bins_after", C.c_int32 * 4)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_uint32), ("ms_total", C.c_float)]

_lib = None


def load():
    """Load the shared library (built by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is the product; there is no fallback path)")
    L = C.CDLL(LIB_PATH)
    vp, sz = C.c_void_p, C.c_size_t
    L.ldvb_abi_version.restype = C.c_int
    L.ldvb_strerror.restype = C.c_char_p
    L.ldvb_strerror.argtypes = [C.c_int]
    L.ldvb_last_error.restype = C.c_char_p
    L.ldvb_last_error.argtypes = [vp]
    L.ldvb_config_default.argtypes = [C.POINTER(Config)]
    L.ldvb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.ldvb_destroy.argtypes = [vp]
    L.ldvb_push.argtypes = [vp, vp, sz]
    L.ldvb_pull.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.ldvb_flush.argtypes = [vp]
    L.ldvb_host_register.argtypes = [vp, sz]
    L.ldvb_host_unregister.argtypes = [vp]
    L.ldvb_process_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.ldvb_get_meas.argtypes = [vp, C.POINTER(Meas)]
    L.ldvb_tap.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
    L.ldvb_table.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
    L.ldvb_host_table.argtypes = [C.POINTER(Config), C.c_int, vp, sz, C.POINTER(sz)]
    L.ldvb_state_size.restype = sz
    L.ldvb_state_size.argtypes = [vp]
    L.ldvb_get_state.argtypes = [vp, vp, sz]
    L.ldvb_set_state.argtypes = [vp, vp, sz]
    L.ldvb_get_rx_state.argtypes = [vp, vp]
    L.ldvb_set_rx_state.argtypes = [vp, vp]
    L.ldvb_fir_cf32.argtypes = [C.c_int, vp, sz, vp, C.c_uint32, C.c_uint32, vp, sz, C.POINTER(sz)]
    L.ldvb_deint_rs.argtypes = [C.c_int, vp, sz, vp, sz, C.POINTER(sz), vp]
    L.ldvb_rs_decode.argtypes = [C.c_int, vp, sz, vp, vp]
    L.ldvb_reset.argtypes = [vp]
    L.ldvb_set_stream.argtypes = [vp, vp]
    L.ldvb_profile.argtypes = [vp, C.c_int]
    L.ldvb_get_profile.argtypes = [vp, C.POINTER(KernelStat), C.c_int, C.POINTER(C.c_int)]
    L.ldvb_pull_cnr.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.ldvb_pull_spectrum.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.ldvb_pull_vber.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.ldvb_edge_size.restype = sz
    L.ldvb_shard_min_halo.restype = sz
    L.ldvb_shard_min_halo.argtypes = [vp]
    L.ldvb_shard_detect.argtypes = [vp, C.POINTER(Shard)]
    L.ldvb_shard_front.argtypes = [vp, C.POINTER(Shard)]
    L.ldvb_shard_back.argtypes = [vp, vp, vp, sz, C.POINTER(sz), vp]
    L.ldvb_ring_unique_id.argtypes = [vp, sz]
    L.ldvb_ring_init.argtypes = [vp, vp, vp, C.c_int, C.c_int]
    L.ldvb_ring_round.argtypes = [vp, C.POINTER(Shard), vp, sz, C.POINTER(sz)]
    L.ldvb_ring_flush.argtypes = [vp]
    L.ldvb_ring_stats.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    L.ldvb_ring_destroy.argtypes = [vp]
    _lib = L
    return L


def default_config(**kw) -> Config:
    cfg = Config()
    load().ldvb_config_default(C.byref(cfg))
    for k, v in kw.items():
        if k == "fmt":
            cfg.input_format = FMT[v]
        elif k == "cstln":
            cfg.constellation = CSTLN[v]
        elif k == "fec":
            cfg.fec = FEC[v] if isinstance(v, str) else v
        elif k == "sampler":
            cfg.sampler = SAMPLER[v] if isinstance(v, str) else v
        elif k == "sub_batch":
            cfg.push_sub_batch = int(v)
        elif k in ("resample", "viterbi", "hard_metric", "fastlock", "allow_drift", "keep_taps", "vber", "hs", "async_push"):
            setattr(cfg, k, int(v))
        else:
            setattr(cfg, k, v)
    return cfg


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def host_register(arr: np.ndarray) -> None:
    """ldvb_host_register: page-locks a host array the caller owns (what the runnable does with its pipebuf)."""
    rc = load().ldvb_host_register(_p(arr), arr.nbytes)
    if rc:
        raise LdvbError(rc, "ldvb_host_register", load().ldvb_strerror(rc).decode())


def host_unregister(arr: np.ndarray) -> None:
    rc = load().ldvb_host_unregister(_p(arr))
    if rc:
        raise LdvbError(rc, "ldvb_host_unregister", load().ldvb_strerror(rc).decode())


def ring_unique_id() -> bytes:
    """ncclGetUniqueId through the library (one rank calls it, the bytes go to the others)."""
    b = np.zeros(128, np.uint8)
    rc = load().ldvb_ring_unique_id(_p(b), b.size)
    if rc:
        raise LdvbError(rc, "ldvb_ring_unique_id", load().ldvb_strerror(rc).decode())
    return b.tobytes()


def host_table(cfg: Config, name: str) -> np.ndarray:
    """Constant table built by the library's host code; needs no GPU."""
    L = load()
    n = C.c_size_t(0)
    rc = L.ldvb_host_table(C.byref(cfg), TABLE[name], None, 0, C.byref(n))
    if rc:
        raise LdvbError(rc, "ldvb_host_table", L.ldvb_strerror(rc).decode())
    out = np.empty(n.value, np.uint8)
    rc = L.ldvb_host_table(C.byref(cfg), TABLE[name], _p(out), out.size, C.byref(n))
    if rc:
        raise LdvbError(rc, "ldvb_host_table", L.ldvb_strerror(rc).decode())
    return out


class Receiver:
    """Host-side mirror of the chain of reference runnables replaced by one handle.

    push()/pull() move HOST buffers like a runnable's run() would between two
    pipebufs; process_device() works on device pointers (torch tensors)."""

    def __init__(self, cfg: Config | None = None, **kw):
        self.L = load()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self.h = C.c_void_p()
        rc = self.L.ldvb_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise LdvbError(rc, "ldvb_create", self.L.ldvb_strerror(rc).decode())
        self.fmt = {v: k for k, v in FMT.items()}[self.cfg.input_format]

    def close(self):
        if self.h:
            self.L.ldvb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, where):
        if rc != 0:
            raise LdvbError(rc, where, self.L.ldvb_last_error(self.h).decode() or
                            self.L.ldvb_strerror(rc).decode())

    def push(self, iq: np.ndarray):
        iq = np.ascontiguousarray(iq, dtype=FMT_DTYPE[self.fmt]).reshape(-1)
        self._ck(self.L.ldvb_push(self.h, _p(iq), iq.size // 2), "ldvb_push")

    def push_ptr(self, host_ptr: int, n_samples: int):
        self._ck(self.L.ldvb_push(self.h, C.c_void_p(host_ptr), n_samples), "ldvb_push")

    def pull(self, max_packets: int = 1 << 20) -> np.ndarray:
        out = np.empty((max_packets, 188), np.uint8)
        n = C.c_size_t(0)
        self._ck(self.L.ldvb_pull(self.h, _p(out), max_packets, C.byref(n)), "ldvb_pull")
        return out[:n.value]

    def pull_ptr(self, host_ptr: int, cap_packets: int) -> int:
        """ldvb_pull into caller memory (address of cap_packets * 188 bytes); returns the packet count."""
        n = C.c_size_t(0)
        self._ck(self.L.ldvb_pull(self.h, C.c_void_p(host_ptr), cap_packets, C.byref(n)), "ldvb_pull")
        return n.value

    def flush(self) -> None:
        """Waits for everything pushed so far (async_push); a no-op otherwise."""
        self._ck(self.L.ldvb_flush(self.h), "ldvb_flush")

    def pull_all(self) -> np.ndarray:
        self.flush()
        parts = []
        while True:
            p = self.pull(1 << 16)
            if p.shape[0] == 0:
                break
            parts.append(p)
        if len(parts) == 1:
            return parts[0]
        return np.concatenate(parts) if parts else np.zeros((0, 188), np.uint8)

    def process_device(self, iq_ptr: int, n_samples: int, ts_ptr: int, cap_packets: int) -> int:
        n = C.c_size_t(0)
        self._ck(self.L.ldvb_process_device(self.h, C.c_void_p(iq_ptr), n_samples,
                                            C.c_void_p(ts_ptr), cap_packets, C.byref(n)),
                 "ldvb_process_device")
        return n.value

    def reset(self):
        self._ck(self.L.ldvb_reset(self.h), "ldvb_reset")

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.ldvb_set_stream(self.h, C.c_void_p(cuda_stream)), "ldvb_set_stream")

    def profile(self, enable: bool):
        self._ck(self.L.ldvb_profile(self.h, int(enable)), "ldvb_profile")

    def get_profile(self) -> dict:
        arr = (KernelStat * 64)()
        n = C.c_int(0)
        self._ck(self.L.ldvb_get_profile(self.h, arr, 64, C.byref(n)), "ldvb_get_profile")
        return {arr[i].name.decode(): {"launches": arr[i].launches, "ms_total": arr[i].ms_total}
                for i in range(n.value)}

    def meas(self) -> dict:
        m = Meas()
        self._ck(self.L.ldvb_get_meas(self.h, C.byref(m)), "ldvb_get_meas")
        return m.asdict()

    def tap(self, name: str) -> np.ndarray:
        which = TAP[name]
        n = C.c_size_t(0)
        self._ck(self.L.ldvb_tap(self.h, which, None, 0, C.byref(n)), "ldvb_tap")
        out = np.empty(n.value, np.uint8)
        if n.value:
            self._ck(self.L.ldvb_tap(self.h, which, _p(out), out.size, C.byref(n)), "ldvb_tap")
        return out

    def table(self, name: str) -> np.ndarray:
        which = TABLE[name]
        n = C.c_size_t(0)
        self._ck(self.L.ldvb_table(self.h, which, None, 0, C.byref(n)), "ldvb_table")
        out = np.empty(n.value, np.uint8)
        self._ck(self.L.ldvb_table(self.h, which, _p(out), out.size, C.byref(n)), "ldvb_table")
        return out

    def rx_state(self) -> np.ndarray:
        w = np.zeros(22, np.uint32)
        self._ck(self.L.ldvb_get_rx_state(self.h, _p(w)), "ldvb_get_rx_state")
        return w

    def set_rx_state(self, w):
        w = np.ascontiguousarray(w, np.uint32)
        self._ck(self.L.ldvb_set_rx_state(self.h, _p(w)), "ldvb_set_rx_state")

    def pull_cnr(self, cap: int = 4096) -> np.ndarray:
        """p_cnr: one C/N value (dB) per second of signal (--cnr)."""
        out = np.zeros(cap, np.float32); n = C.c_size_t(0)
        self._ck(self.L.ldvb_pull_cnr(self.h, _p(out), cap, C.byref(n)), "ldvb_pull_cnr")
        return out[: n.value]

    def pull_vber(self, cap: int = 4096) -> np.ndarray:
        """p_vber: rate_estimator over the RS decoder's counts (needs vber=True)."""
        out = np.zeros(cap, np.float32); n = C.c_size_t(0)
        self._ck(self.L.ldvb_pull_vber(self.h, _p(out), cap, C.byref(n)), "ldvb_pull_vber")
        return out[: n.value]

    def pull_spectrum(self, cap_rows: int = 1024) -> np.ndarray:
        """p_spectrum: float[1024] rows (dB, fft-shifted), one per second of signal."""
        out = np.zeros((cap_rows, 1024), np.float32); n = C.c_size_t(0)
        self._ck(self.L.ldvb_pull_spectrum(self.h, _p(out), cap_rows, C.byref(n)), "ldvb_pull_spectrum")
        return out[: n.value]

    # ---- time sharding (one stream over several handles)
    def edge_size(self) -> int:
        return int(self.L.ldvb_edge_size())

    def shard_min_halo(self) -> int:
        return int(self.L.ldvb_shard_min_halo(self.h))

    def shard(self, iq_ptr: int, abs_raw0: int, n_halo: int, n_chunk: int, n_halo_next: int, last: bool = False,
              bins_before=(-1, -1, -1, -1)) -> Shard:
        s = Shard()
        s.iq_dev = iq_ptr; s.abs_raw0 = abs_raw0; s.n_halo = n_halo; s.n_chunk = n_chunk
        s.n_halo_next = n_halo_next; s.last = 1 if last else 0
        for i in range(4):
            s.bins_before[i] = int(bins_before[i]); s.bins_after[i] = -1
        return s

    def shard_detect(self, s: Shard):
        """Fills s.bins_after; returns it as a tuple."""
        self._ck(self.L.ldvb_shard_detect(self.h, C.byref(s)), "ldvb_shard_detect")
        return tuple(int(v) for v in s.bins_after)

    def shard_front(self, s: Shard):
        self._ck(self.L.ldvb_shard_front(self.h, C.byref(s)), "ldvb_shard_front")

    def shard_back(self, edge_in, ts_ptr: int, cap_packets: int, edge_out) -> int:
        """edge_in / edge_out: uint8 numpy arrays of edge_size() bytes, or None."""
        n = C.c_size_t(0)
        self._ck(self.L.ldvb_shard_back(self.h, _p(edge_in) if edge_in is not None else None, ts_ptr, cap_packets,
                                        C.byref(n), _p(edge_out) if edge_out is not None else None), "ldvb_shard_back")
        return int(n.value)

    # ---- the ring inside the library (NCCL send/recv between neighbouring ranks)
    def ring_init(self, id_early: bytes, id_edge: bytes, rank: int, nranks: int):
        a = np.frombuffer(bytes(id_early), np.uint8).copy(); b = np.frombuffer(bytes(id_edge), np.uint8).copy()
        self._ck(self.L.ldvb_ring_init(self.h, _p(a), _p(b), rank, nranks), "ldvb_ring_init")

    def ring_round(self, s: Shard, ts_ptr: int, cap_packets: int) -> int:
        n = C.c_size_t(0)
        self._ck(self.L.ldvb_ring_round(self.h, C.byref(s), ts_ptr, cap_packets, C.byref(n)), "ldvb_ring_round")
        return int(n.value)

    def ring_flush(self):
        self._ck(self.L.ldvb_ring_flush(self.h), "ldvb_ring_flush")

    def ring_stats(self, reset: bool = False) -> dict:
        v = (C.c_double * 4)()
        self._ck(self.L.ldvb_ring_stats(self.h, v, 1 if reset else 0), "ldvb_ring_stats")
        return dict(zip(("early", "front", "wait_edge", "back"), (float(x) for x in v)))

    def ring_destroy(self):
        self._ck(self.L.ldvb_ring_destroy(self.h), "ldvb_ring_destroy")

    def get_state(self) -> np.ndarray:
        n = self.L.ldvb_state_size(self.h)
        b = np.zeros(n, np.uint8)
        self._ck(self.L.ldvb_get_state(self.h, _p(b), n), "ldvb_get_state")
        return b

    def set_state(self, b):
        b = np.ascontiguousarray(b, np.uint8)
        self._ck(self.L.ldvb_set_state(self.h, _p(b), b.size), "ldvb_set_state")


def fir_cf32(x: np.ndarray, taps_cplx: np.ndarray, decim: int = 1, device: int = 0) -> np.ndarray:
    L = load()
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    t = np.ascontiguousarray(taps_cplx, np.float32).reshape(-1)
    n_in, nt = x.size // 2, t.size // 2
    out = np.empty(2 * (n_in // decim + 1), np.float32)
    n = C.c_size_t(0)
    rc = L.ldvb_fir_cf32(device, _p(x), n_in, _p(t), nt, decim, _p(out), out.size // 2, C.byref(n))
    if rc:
        raise LdvbError(rc, "ldvb_fir_cf32", L.ldvb_strerror(rc).decode())
    return out[:2 * n.value]


def rs_decode(packets204: np.ndarray, device: int = 0):
    L = load()
    p = np.ascontiguousarray(packets204, np.uint8).reshape(-1, 204)
    out = np.empty((p.shape[0], 188), np.uint8)
    flags = np.zeros((p.shape[0], 2), np.int32)
    rc = L.ldvb_rs_decode(device, _p(p), p.shape[0], _p(out), _p(flags))
    if rc:
        raise LdvbError(rc, "ldvb_rs_decode", L.ldvb_strerror(rc).decode())
    return out, flags


def deint_rs(mpegbytes: np.ndarray, device: int = 0):
    L = load()
    m = np.ascontiguousarray(mpegbytes, np.uint8).reshape(-1)
    cap = m.size // 204 + 1
    out = np.empty((cap, 188), np.uint8)
    flags = np.zeros((cap, 2), np.int32)
    n = C.c_size_t(0)
    rc = L.ldvb_deint_rs(device, _p(m), m.size, _p(out), cap, C.byref(n), _p(flags))
    if rc:
        raise LdvbError(rc, "ldvb_deint_rs", L.ldvb_strerror(rc).decode())
    return out[:n.value], flags[:n.value]
