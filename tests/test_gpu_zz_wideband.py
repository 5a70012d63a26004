"""BASELINE.json configs[4], reading 5b of SURVEY 8(d), through the whole CUDA chain: a carrier oversampled 120x,
`--resample` = 313-tap low-pass with decimation 30 in k_frontend, then the 4 samples/symbol receiver."""
import numpy as np
import pytest

from tests import vectors as V
from tests.test_gpu_parity import assert_prefix, run_product

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_wideband_resample_chain(product, oracle, mode):
    P, O = product, oracle
    raw = V.ref_iq(120, ratio="120", fmt="f32")
    kw = dict(fmt="f32", resample=True, Fs=240e6)
    ref = O.Chain(O.Config(**kw)).run(raw)
    got = run_product(P, raw, rx_mode=P.RX_EXACT if mode == "exact" else P.RX_FAST, **kw)
    assert_prefix(got["pp"], ref["pp"], "preprocessed IQ (313 taps, decimation 30)")
    if mode == "exact":
        assert_prefix(got["symbols"], ref["symbols"], "soft symbols")
        assert_prefix(got["bytes"], ref["bytes"], "deconvolved bytes", slack=8)
    assert_prefix(got["ts"], ref["ts"], "TS")
    assert len(ref["ts"]) >= 40
    # 4 samples per symbol after the decimation: the handle runs the exact receiver whatever is asked
    # (include/leandvb_b200.h, "Receiver scheduling mode")
    assert got["meas"]["seams_total"] == 0
