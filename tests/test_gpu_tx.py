"""Transmit chain on the GPU (include/leandvb_b200_tx.h) against the oracle and the reference:
every stream bit-exact (bytes, symbols and the cf32 output: explicit _rn arithmetic in the
reference's order, host-built taps).  Run with `pytest -m gpu` on a B200."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import vectors as V
from tests.conftest import ROOT
from tests.tx_cases import TX_CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("case", TX_CASES, ids=[c[0] for c in TX_CASES])
def test_tx_one_shot_equals_oracle_and_reference_digest(product, oracle, case):
    name, npk, cst, cr, ratio, power, agc, rolloff = case
    ts = V.ts_packets(npk)
    want = oracle.tx_chain(ts, cst, cr, ratio, power, agc, rolloff)
    tx = product.Transmitter(cstln=cst, fec=cr, ratio=ratio, power=power, agc=agc, rolloff=rolloff, max_packets=npk,
                             keep_taps=True)
    got = tx.push(ts)
    assert np.array_equal(tx.tap("rspackets").reshape(-1, 204), oracle.tx_rs_packets(ts))
    assert np.array_equal(tx.tap("mpegbytes"), want["mpegbytes"])
    assert np.array_equal(tx.tap("symbols"), want["symbols"])
    assert got.size == want["iq"].size
    assert np.array_equal(_bits(got), _bits(want["iq"]))
    kat = json.load(open(os.path.join(GOLDEN, "tx_kat.json")))[name]     # made by the reference binary
    assert got.size // 2 == kat["samples"] and hashlib.sha256(got.tobytes()).hexdigest() == kat["sha256"]
    tx.close()


@pytest.mark.parametrize("case", [TX_CASES[0], TX_CASES[2], TX_CASES[3], TX_CASES[7]], ids=lambda c: c[0])
@pytest.mark.parametrize("steps", [(1, 2, 3, 5, 8, 13, 40, 1, 1, 200), (12, 1, 64), (7,) * 40])
def test_tx_streaming_equals_one_shot(product, oracle, case, steps):
    """Packets pushed in ragged batches (including batches smaller than the interleaver depth,
    the filter history and one AGC chunk): the concatenated output equals the one-shot output."""
    name, npk, cst, cr, ratio, power, agc, rolloff = case
    ts = V.ts_packets(npk)
    want = oracle.tx_chain(ts, cst, cr, ratio, power, agc, rolloff)["iq"]
    tx = product.Transmitter(cstln=cst, fec=cr, ratio=ratio, power=power, agc=agc, rolloff=rolloff, max_packets=256)
    parts, at = [], 0
    for s in steps:
        if at >= npk:
            break
        parts.append(tx.push(ts[at: at + s]).copy())
        at += s
    if at < npk:
        parts.append(tx.push(ts[at:]).copy())
    parts.append(tx.push(ts[:0]).copy())            # empty push: no progress, no error
    got = np.concatenate(parts)
    assert got.size == want.size
    assert np.array_equal(_bits(got), _bits(want))
    tx.reset()                                     # a reset handle starts the same stream again
    again = tx.push(ts[:64])
    assert np.array_equal(_bits(again), _bits(want[: again.size])) and again.size > 0
    tx.close()


def test_fir_resampler_stage_equals_oracle(product, oracle):
    """fir_resampler<cf32,float> alone (dsp.h:290-364) on seeded noise, several interpolation
    factors, and the starved cases (fewer than ncoeffs inputs: no output)."""
    rng = np.random.default_rng(7)
    for interp, n_in in ((6, 5000), (2, 777), (5, 51), (5, 50), (120, 1300), (3, 0)):
        taps = oracle.tx_taps(interp, 0.35, 10.0, "12.5")
        x = (rng.standard_normal(2 * n_in) * 50).astype(np.float32)
        want = oracle.tx_resample(x, taps, interp)
        cplx = np.zeros(2 * taps.size, np.float32)
        cplx[0::2] = taps * np.float32(1.0)          # set_freq(0): cosf(0) = 1, sinf(0) = 0 (dsp.h:352-361)
        cplx[1::2] = taps * np.float32(0.0)
        got = product.fir_resampler_cf32(x, cplx, interp) if n_in else np.zeros(0, np.float32)
        assert got.size == want.size, (interp, n_in)
        assert np.array_equal(_bits(got), _bits(want)), (interp, n_in)


def test_device_resident_tx_into_rx_loopback(product, oracle):
    """leantsgen-style packets generated in HBM -> transmit chain -> receive chain, all device
    resident: the decoded TS is the transmitted numbered stream (BER 0), and the
    IQ equals the oracle's."""
    import torch
    P = product
    npk = 2000
    dev = torch.device("cuda", 0)
    tx = P.Transmitter(ratio="6/5", power="37.5", agc=True, max_packets=npk)
    ts_dev = torch.empty(npk * 188, dtype=torch.uint8, device=dev)
    tx.tsgen_device(0, npk, ts_dev.data_ptr())
    cap = tx.max_samples(npk)
    iq_dev = torch.empty(2 * cap, dtype=torch.float32, device=dev)
    n = tx.process_device(ts_dev.data_ptr(), npk, iq_dev.data_ptr(), cap)
    torch.cuda.synchronize()
    assert np.array_equal(ts_dev.cpu().numpy().reshape(-1, 188), V.ts_packets(npk))
    want = oracle.tx_chain(V.ts_packets(npk), "QPSK", "1/2", "6/5", "37.5", True)["iq"]
    got = iq_dev[: 2 * n].cpu().numpy()
    assert got.size == want.size and np.array_equal(_bits(got), _bits(want))
    rx = P.Receiver(fmt="f32", resample=True, rx_mode=P.RX_FAST, max_batch=n)
    out_dev = torch.empty((npk + 64) * 188, dtype=torch.uint8, device=dev)
    k = rx.process_device(iq_dev.data_ptr(), n, out_dev.data_ptr(), npk + 64)
    ts = out_dev[: k * 188].cpu().numpy().reshape(-1, 188)
    ctr = (ts[:, 1].astype(int) << 16) | (ts[:, 2].astype(int) << 8) | ts[:, 3]
    assert k > npk - 100      # 11 packets stay in the interleaver, ~50 go into acquisition
    assert np.array_equal(ts[3:], V.ts_packets(k - 3, int(ctr[3])))
    rx.close(); tx.close()


def test_tx_rejects_unsuitable_code_rate(product):
    """leandvbtx fail()s with "Code rate not suitable for this constellation" (dvb.h:582-584) or
    "Code rate not supported with APSK16" (dvb.h:59); the handle refuses the same combinations."""
    from tests.tx_cases import TX_REJECTED
    for cst, cr in TX_REJECTED:
        with pytest.raises(product.LdvbError):
            product.Transmitter(cstln=cst, fec=cr, ratio="2", power="37.5", agc=False, max_packets=16)
