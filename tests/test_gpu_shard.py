"""Time-sharded mode on ONE GPU: the stream is cut into chunks that go
through ldvb_shard_detect / ldvb_shard_front / ldvb_shard_back on alternating handles, the
EDGE blob being the only thing that passes from one chunk to the next.  The TS must be the
oracle's, bit for bit.  (The multi-process transport around the same calls is covered by
tests/test_shard_ring_cpu.py on gloo and by `bench.py --gpus N` on NCCL.)"""
import numpy as np
import pytest

from tests import vectors as V
from tests.test_gpu_parity import _freq_shift

pytestmark = pytest.mark.gpu


def run_sharded(P, raw, n_chunks, n_engines, halo_units=None, unit=4096, **kw):
    import torch
    from leansdr_b200 import shard as S
    n = raw.size // 2
    dev = torch.device("cuda", 0)
    iq = torch.from_numpy(raw).to(dev)
    bps = raw.itemsize * 2
    cap = n // 1500 + 64
    rxs, engines, ts_bufs = [], [], []
    for _ in range(n_engines):
        rx = P.Receiver(rx_mode=P.RX_FAST, max_batch=n, **kw)
        ts = torch.empty(cap * 188, dtype=torch.uint8, device=dev)
        rxs.append(rx); ts_bufs.append(ts)
        engines.append(S.GpuEngine(rx, ts.data_ptr(), cap))
    # unit: a multiple of lcm(4096, 128 * decimation); 4096 for decimation 1 and 2
    halo = -(-rxs[0].shard_min_halo() // unit) * unit if halo_units is None else halo_units * unit
    chunks = S.plan_stream(n, n_chunks, unit, halo)
    out = []
    bins, edge = (-1, -1, -1, -1), None
    for ch in chunks:
        e = engines[ch.index % n_engines]
        after = e.detect(ch, iq.data_ptr() + ch.abs_raw0 * bps, bins)
        e.front()
        npk, edge = e.back(edge, not ch.last)
        out.append(ts_bufs[ch.index % n_engines][: npk * 188].cpu().numpy().reshape(-1, 188))
        bins = after
    meas = [rx.meas() for rx in rxs]
    for rx in rxs:
        rx.close()
    return np.concatenate(out), meas, chunks


def _tail(raw, chunks):
    """Packets that may be missing at the end: the samples that plan_stream() leaves out
    (less than one alignment unit per chunk), at ~1958 samples per packet and more, + 2."""
    unused = raw.size // 2 - (chunks[-1].start + chunks[-1].n_chunk)
    return 2 + -(-unused // 1900)


def check_ts(got, want, lost_tail=2):
    n = min(len(got), len(want))
    assert n > 0.9 * len(want)
    assert np.array_equal(got[:n], want[:n]), f"first differing packet {int(np.nonzero((got[:n] != want[:n]).any(axis=1))[0][0])} of {n}"
    assert -lost_tail <= len(got) - len(want) <= 1, (len(got), len(want))


CASES = [
    # name, receiver kw, generator kw, packets, chunks, engines
    ("f32-resample-anf1-3x2", dict(fmt="f32", resample=True), {}, 7000, 3, 2),
    ("u8-anf0-noise-4x1", dict(fmt="u8", anf=0), dict(noise_db=22), 1600, 4, 1),
    ("f32-anf2-derot-2x2", dict(fmt="f32", anf=2, Fderot=20000.0), {}, 5000, 2, 2),
    ("f32-decim2-3x3", dict(fmt="f32", anf=0, decim=2, Fs=4.8e6, float_scale=0.5), dict(ratio="12/5", power=43.5), 1200, 3, 3),
    ("viterbi-2x2", dict(fmt="f32", viterbi=True), dict(noise_db=25), 500, 2, 2),
]


@pytest.mark.parametrize("name,kw,gkw,npk,nch,neng", CASES, ids=[c[0] for c in CASES])
def test_time_sharded_ts_bit_exact(product, oracle, name, kw, gkw, npk, nch, neng):
    P, O = product, oracle
    raw = V.ref_iq(npk, fmt=kw["fmt"], **gkw)
    want = O.Chain(O.Config(**kw)).run(raw)["ts"]
    got, meas, chunks = run_sharded(P, raw, nch, neng, **kw)
    check_ts(got, want, lost_tail=_tail(raw, chunks))
    assert sum(m["seams_total"] for m in meas) > nch


def test_time_sharded_wideband_resample_chain(product, oracle):
    """BASELINE.json configs[4], reading 5b: a carrier oversampled 120x (Fs/Fm = 120), `--resample` = 313-tap low-pass
    with decimation 30, then the 4 samples/symbol receiver -- time-sharded over three handles.  The waveform arrives
    ~11x below the nominal level: the first chunk settles its AGC serially, the later chunks seed their warm-ups with
    the measured power; alignment unit lcm(4096, 128 * 30) = 61440 samples."""
    P, O = product, oracle
    raw = V.ref_iq(120, ratio="120", fmt="f32")
    kw = dict(fmt="f32", resample=True, Fs=240e6)
    want = O.Chain(O.Config(**kw)).run(raw)["ts"]
    got, meas, chunks = run_sharded(P, raw, 3, 3, unit=61440, **kw)
    assert len(want) >= 40
    check_ts(got, want, lost_tail=_tail(raw, chunks))
    assert sum(m["seams_total"] for m in meas) > 3 and sum(m["settle_passes"] for m in meas) >= 1


def test_time_sharded_cold_start_with_carrier_offset(product, oracle):
    """Every handle starts with freqw = 0 while the carrier sits 1.5e-3 cycles/sample off:
    the warm-ups of all spans (including the one in the halo) have to be re-seeded."""
    P, O = product, oracle
    raw = _freq_shift(V.ref_iq(2400, fmt="f32"), 1.5e-3)
    kw = dict(fmt="f32", resample=True)
    want = O.Chain(O.Config(**kw)).run(raw)["ts"]
    got, meas, chunks = run_sharded(P, raw, 3, 3, **kw)
    check_ts(got, want, lost_tail=_tail(raw, chunks))


def test_time_sharded_seam_repair_path(product, oracle):
    """One warm-up chunk is too short for the loops to converge: seams fail verification (the
    one between chunks too) and are repaired from the imported loop state."""
    P, O = product, oracle
    raw = V.ref_iq(1600, fmt="f32", noise_db=22)
    kw = dict(fmt="f32", resample=True)
    want = O.Chain(O.Config(**kw)).run(raw)["ts"]
    got, meas, chunks = run_sharded(P, raw, 4, 2, warmup_chunks=1, span_chunks=4, **kw)
    check_ts(got, want, lost_tail=_tail(raw, chunks))
    assert sum(m["seams_repaired"] for m in meas) > 0


def test_time_sharded_rejects_bad_geometry(product):
    import torch
    P = product
    rx = P.Receiver(fmt="f32", rx_mode=P.RX_FAST, max_batch=1 << 20)
    buf = torch.zeros(2 << 20, dtype=torch.float32, device="cuda:0")
    h = rx.shard_min_halo()
    assert h % 4096 == 0 and h >= 4096 * 5
    with pytest.raises(P.LdvbError):      # misaligned chunk
        rx.shard_front(rx.shard(buf.data_ptr(), 0, 0, 4096 * 100 + 128, 0, True))
    with pytest.raises(P.LdvbError):      # halo too short
        rx.shard_front(rx.shard(buf.data_ptr(), 4096 * 50, 4096, 4096 * 100, 0, True))
    with pytest.raises(P.LdvbError):      # back without front
        rx.shard_back(None, buf.data_ptr(), 16, None)
    ex = P.Receiver(fmt="f32", rx_mode=P.RX_EXACT, max_batch=1 << 20)
    with pytest.raises(P.LdvbError):      # exact mode has no speculative front stage
        ex.shard_front(ex.shard(buf.data_ptr(), 0, 0, 4096 * 100, 0, True))
    rx.close(); ex.close()
