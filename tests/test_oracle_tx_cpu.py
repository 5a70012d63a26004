"""The transmit-chain oracle (oracle/dvbs_tx_oracle.c) pinned against the reference:
  * the UNMODIFIED reference binary oracle/_ref/leandvbtx on the same numbered packets
    (when oracle/_ref is present: the dev container and the GPU box),
  * committed SHA-256 digests of the reference's output (tests/golden/tx_kat.json, made by
    tests/golden/make_golden.py from the reference binary).
Also: the product's host-side tap builder (no device needed) equals the oracle's."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests import vectors as V
from tests.conftest import ROOT
from tests.tx_cases import TX_CASES

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ref_tx(O, npk, cst, cr, ratio, power, agc, rolloff):
    args = [O.ref_bin("leandvbtx"), "--const", cst, "--cr", cr, "-f", ratio, "--power", power, "--roll-off", str(rolloff)]
    if agc:
        args.append("--agc")
    out = subprocess.run(args, input=V.ts_packets(npk).tobytes(), stdout=subprocess.PIPE, check=True).stdout
    return np.frombuffer(out, dtype=np.float32)


@pytest.mark.parametrize("case", TX_CASES, ids=[c[0] for c in TX_CASES])
def test_tx_oracle_equals_reference_digest(oracle, case):
    name, npk, cst, cr, ratio, power, agc, rolloff = case
    kat = json.load(open(os.path.join(GOLDEN, "tx_kat.json")))[name]
    got = oracle.tx_chain(V.ts_packets(npk), cst, cr, ratio, power, agc, rolloff)["iq"]
    assert got.size // 2 == kat["samples"]
    assert hashlib.sha256(got.tobytes()).hexdigest() == kat["sha256"]


@pytest.mark.skipif(not V.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", TX_CASES[:4], ids=[c[0] for c in TX_CASES[:4]])
def test_tx_oracle_equals_reference_binary(oracle, case):
    name, npk, cst, cr, ratio, power, agc, rolloff = case
    want = _ref_tx(oracle, npk, cst, cr, ratio, power, agc, rolloff)
    got = oracle.tx_chain(V.ts_packets(npk), cst, cr, ratio, power, agc, rolloff)["iq"]
    assert got.size == want.size
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_tx_loopback_through_rx_oracle(oracle):
    """What the transmit oracle makes, the receive oracle decodes back to the numbered packets."""
    O = oracle
    iq = O.tx_chain(V.ts_packets(160), "QPSK", "1/2", "6/5", "37.5", True)["iq"]
    ts = O.Chain(O.Config(fmt="f32")).run(iq)["ts"]
    ctr = (ts[:, 1].astype(int) << 16) | (ts[:, 2].astype(int) << 8) | ts[:, 3]
    assert len(ts) > 100
    assert np.array_equal(ts[3:], V.ts_packets(len(ts) - 3, int(ctr[3])))


@pytest.mark.parametrize("ratio,power,rolloff", [("6/5", "37.5", 0.35), ("2", "0", 0.35), ("5", "37.5", 0.2), ("120", "20", 0.35)])
def test_product_host_taps_equal_oracle(oracle, product, ratio, power, rolloff):
    cfg = product.tx_config(ratio=ratio, power=power, rolloff=rolloff)
    got = product.host_taps(cfg)
    want = oracle.tx_taps(cfg.interp, rolloff, 10.0, power)
    assert got.size == want.size == ((int(cfg.interp * 10.0) + 1) | 1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
