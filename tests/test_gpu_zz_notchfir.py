"""The two auto_notch kernels side by side: the warp-specialised one of k_notchfir.cu (the default: plain notch, and notch +
fir_filter fused on the store path so that the notched stream never reaches HBM) and the lane-per-segment one of k_notch.cu
(LDVB_NOTCH_V2=0, what time-sharded handles run).  Same bytes as the oracle in every stream; batches chosen so that segment
boundaries, the carried FIR history between batches and the telemetry blocks (cnr_fft / spectrum read the notched stream)
are all exercised."""
import numpy as np
import pytest

from tests import vectors as V
from tests.test_gpu_parity import assert_prefix, run_product

pytestmark = pytest.mark.gpu

CASES = [
    ("f32-resample-fused", dict(fmt="f32", resample=True), {}, 1500, None),
    ("f32-resample-fused-batches", dict(fmt="f32", resample=True), {}, 1500, 700001),
    ("u8-plain-notch", dict(fmt="u8"), {}, 600, None),
    ("f32-anf2-derot-unfused", dict(fmt="f32", anf=2, Fderot=20000.0), {}, 600, 300007),
]


@pytest.mark.parametrize("v2", ["1", "0"])
@pytest.mark.parametrize("name,kw,gkw,npk,batch", CASES, ids=[c[0] for c in CASES])
def test_notchfir_kernel_every_stream_bit_exact(product, oracle, monkeypatch, name, kw, gkw, npk, batch, v2):
    P, O = product, oracle
    monkeypatch.setenv("LDVB_NOTCH_V2", v2)            # read by ldvb_create
    raw = V.ref_iq(npk, fmt=kw["fmt"], **gkw)
    ref = O.Chain(O.Config(**kw)).run(raw)
    got = run_product(P, raw, n_batch=batch, rx_mode=P.RX_EXACT, **kw)
    assert_prefix(got["pp"], ref["pp"], "preprocessed IQ")
    assert_prefix(got["symbols"], ref["symbols"], "soft symbols")
    assert_prefix(got["ts"], ref["ts"], "TS", slack=188)
    assert len(ref["ts"]) > npk - 80


@pytest.mark.parametrize("v2", ["1", "0"])
@pytest.mark.parametrize("batch", [None, 3_000_001])
def test_notchfir_spectrum_rows(product, oracle, monkeypatch, batch, v2):
    """spectrum (always on, leandvb.cc:333-343) reads the notched stream, which the fused kernel only writes out for the
    blocks that will be measured (known in advance: sdr.h:1362-1370)."""
    from tests.test_gpu_parity import _telemetry
    P, O = product, oracle
    monkeypatch.setenv("LDVB_NOTCH_V2", v2)
    raw = V.ref_iq(4000, fmt="f32")
    kw = dict(fmt="f32", resample=True)
    ref = O.Chain(O.Config(**kw)).run(raw)
    _, rows = _telemetry(P, raw, batch or raw.size // 2, **kw)
    assert len(ref["spectrum"]) >= 3 and np.array_equal(rows, ref["spectrum"][: len(rows)]) and len(rows) >= len(ref["spectrum"]) - 1
