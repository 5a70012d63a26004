"""CPU-only checks of the C-ABI library: it loads, exports every symbol declared in
include/leandvb_b200.h, refuses to run without a device (no fallback), and its
host-side table builders reproduce the reference's tables."""
import hashlib
import json
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, have_gpu

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _declared(header, prefix):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, hdr))


def test_exports_match_header(product):
    """Every include/*.h is covered: leandvb_b200.h (receive path) and leandvb_b200_tx.h (transmit chain)."""
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["leandvb_b200.h", "leandvb_b200_tx.h"]
    L = product.load()
    for header, prefix, listed in (("leandvb_b200.h", "ldvb", product.EXPORTS), ("leandvb_b200_tx.h", "ldvbtx", product.TX_EXPORTS)):
        declared = _declared(header, prefix)
        assert declared, "no declarations parsed"
        missing = [f for f in sorted(declared) if not hasattr(L, f)]
        assert not missing, f"declared in include/{header} but not exported: {missing}"
        assert set(listed) == declared


def test_tx_defaults_and_no_device(product):
    cfg = product.tx_config()
    # leandvbtx.cc:69-76 defaults
    assert (cfg.constellation, cfg.fec, cfg.interp, cfg.decim, cfg.agc) == (1, 0, 2, 1, 0)
    assert abs(cfg.rolloff - 0.35) < 1e-6 and abs(cfg.rrc_rej - 10) < 1e-6 and cfg.power_db == b"0"
    if not have_gpu():
        with pytest.raises(product.LdvbError) as e:
            product.Transmitter(cfg)
        assert e.value.code == -4      # LDVB_ENODEV: no CPU fallback


def test_abi_and_defaults(product):
    L = product.load()
    assert L.ldvb_abi_version() == 3
    c = product.default_config()
    # leandvb.cc:88-135 defaults
    assert (c.input_format, c.anf, c.sampler, c.constellation, c.fec) == (0, 1, 1, 1, 0)
    assert abs(c.Fs - 2.4e6) < 1 and abs(c.Fm - 2e6) < 1 and abs(c.rolloff - 0.35) < 1e-6
    assert c.resample == 0 and c.viterbi == 0 and c.fastlock == 0 and abs(c.Finfo - 5) < 1e-6


@pytest.mark.skipif(have_gpu(), reason="checks the no-device behaviour")
def test_create_fails_loudly_without_gpu(product):
    with pytest.raises(product.LdvbError) as e:
        product.Receiver()
    assert e.value.code == -4  # LDVB_ENODEV: there is no CPU fallback


def test_unsupported_config_is_rejected(product):
    import ctypes as C
    L = product.load()
    cfg = product.default_config()
    cfg.abi_version = 99
    h = C.c_void_p()
    assert L.ldvb_create(C.byref(cfg), C.byref(h)) == -1


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_host_tables_match_reference_dumps(product):
    """Tables dumped from the unmodified reference headers (tests/golden/tables.json,
    made by oracle/_ref/ref_tables) vs the library's own host builders."""
    g = json.load(open(os.path.join(GOLDEN, "tables.json")))
    P = product
    c = P.default_config()
    assert _sha(P.host_table(c, "cstln")) == g["cstln_qpsk.bin"]["sha256"]
    assert _sha(P.host_table(P.default_config(cstln="BPSK"), "cstln")) == g["cstln_bpsk.bin"]["sha256"]
    assert _sha(P.host_table(P.default_config(cstln="8PSK"), "cstln")) == g["cstln_8psk.bin"]["sha256"]
    assert _sha(P.host_table(P.default_config(hard_metric=True), "cstln")) == g["cstln_qpsk_hard.bin"]["sha256"]
    assert _sha(P.host_table(c, "trig16")) == g["trig16.f32"]["sha256"]
    assert _sha(P.host_table(c, "derand")) == g["derand.u8"]["sha256"]
    assert _sha(P.host_table(c, "rs_log")) == g["rs_log.u8"]["sha256"]
    assert _sha(P.host_table(c, "rs_exp")[:511]) == g["rs_exp.u8"]["sha256"]
    assert _sha(P.host_table(P.default_config(resample=True), "fir")) == g["lowpass_fs2.4_sr2.f32"]["sha256"]
    assert _sha(P.host_table(P.default_config(resample=True, Fs=9.6e6), "fir")) == g["lowpass_fs9.6_sr2.f32"]["sha256"]
    assert _sha(P.host_table(P.default_config(resample=True, Fs=240e6), "fir")) == g["lowpass_fs240_sr2.f32"]["sha256"]
    assert _sha(P.host_table(P.default_config(sampler="rrc"), "rrc")) == g["rrc_fs2.4_sr2.f32"]["sha256"]
    assert _sha(P.host_table(P.default_config(sampler="rrc", Fs=4e6), "rrc")) == g["rrc_fs4_sr2.f32"]["sha256"]
    assert _sha(P.host_table(c, "vitmap")) == g["vitmap_qpsk12.u8"]["sha256"]
    assert _sha(P.host_table(P.default_config(fec="7/8"), "vitmap")) == g["vitmap_qpsk78.u8"]["sha256"]
    # more filter designs (filtergen.h:45-92 through leandvb.cc:353-378, 437-456)
    for kw, kind, f in ((dict(resample=True, Fs=9.6e6, resample_rej=20.0), "fir", "lowpass_fs9.6_sr2_rej20.f32"),
                        (dict(resample=True, rolloff=0.2), "fir", "lowpass_fs2.4_sr2_ro02.f32"),
                        (dict(resample=True, Fs=55e6, Fm=27.5e6), "fir", "lowpass_fs55_sr27.5.f32"),
                        (dict(sampler="rrc", rolloff=0.2), "rrc", "rrc_fs2.4_sr2_ro02.f32"),
                        (dict(sampler="rrc", Fs=8e6, rrc_rej=5.0), "rrc", "rrc_fs8_sr2_rej5.f32")):
        assert _sha(P.host_table(P.default_config(**kw), kind)) == g[f]["sha256"], f
    # --hs: fast_qpsk_receiver's three look-up tables (sdr.h:1144-1164)
    for name, f in (("hs_polar", "hs_polar.u32"), ("hs_rect", "hs_rect.u16"), ("hs_sincos", "hs_sincos.u16")):
        assert _sha(P.host_table(c, name)) == g[f]["sha256"], name
    # every constellation of cstln_lut<256>::predef (sdr.h:305-311); APSK radii per code rate (dvb.h:45-81)
    for cst, fec, f in (("16APSK", "2/3", "16apsk23"), ("16APSK", "3/4", "16apsk34"), ("16APSK", "5/6", "16apsk56"),
                        ("32APSK", "3/4", "32apsk34"), ("32APSK", "5/6", "32apsk56"), ("64APSKe", "3/4", "64apske"),
                        ("16QAM", "1/2", "16qam"), ("64QAM", "1/2", "64qam"), ("256QAM", "1/2", "256qam")):
        assert _sha(P.host_table(P.default_config(cstln=cst, fec=fec), "cstln")) == g[f"cstln_{f}.bin"]["sha256"], cst
    assert _sha(P.host_table(P.default_config(cstln="16APSK", fec="3/4", hard_metric=True), "cstln")) == \
        g["cstln_16apsk34_hard.bin"]["sha256"]
    for cst, fec, f in (("8PSK", "2/3", "8psk23"), ("16APSK", "3/4", "16apsk34"), ("16QAM", "3/4", "16qam34"),
                        ("64QAM", "4/6", "64qam46"), ("256QAM", "7/8", "256qam78")):
        assert _sha(P.host_table(P.default_config(cstln=cst, fec=fec), "vitmap")) == g[f"vitmap_{f}.u8"]["sha256"], cst
    for fec, f in (("3/4", "34"), ("4/6", "46"), ("5/6", "56"), ("7/8", "78")):
        t = P.host_table(P.default_config(fec=fec), "trellis").reshape(-1, 2).copy()
        t[t[:, 0] == 65, 1] = 0          # `us` of absent branches: not written by the reference
        assert _sha(t) == g[f"trellis_{f}.bin"]["sha256"], fec
    # combinations the reference fail()s on (dvb.h:59, 70)
    for cst, fec in (("16APSK", "1/2"), ("32APSK", "1/2"), ("32APSK", "2/3"), ("16APSK", "7/8")):
        with pytest.raises(P.LdvbError):
            P.host_table(P.default_config(cstln=cst, fec=fec), "cstln")


def test_known_answers(product):
    """Literals the reference asserts or documents."""
    P = product
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    c = P.default_config()
    dec = P.host_table(c, "deconv").view(np.uint64)
    assert int(dec[0]) == int(kat["deconv_fec12"], 16)          # dvb.h:120, 239
    cells = P.host_table(c, "cstln").view(np.int16).reshape(256, 256, 4)
    assert cells[53, 53, :3].tolist() == kat["lookup_53_53"]
    assert cells[10, (-3) & 255, :3].tolist() == kat["lookup_10_m3"]
    for fec, f in (("1/2", "deconv_12.u64"), ("3/4", "deconv_34.u64"), ("7/8", "deconv_78.u64")):
        ref = np.fromfile(os.path.join(GOLDEN, f), dtype=np.uint64)
        pp = int(ref[0])
        got = P.host_table(P.default_config(fec=fec), "deconv").view(np.uint64)
        assert np.array_equal(got, ref[2:2 + pp])
    # Trellis: compare the branches that exist (the reference leaves `us` of absent
    # branches uninitialised, viterbi.h:52-55).
    ref = np.fromfile(os.path.join(GOLDEN, "trellis_12.bin"), dtype=np.uint8).reshape(-1, 2)
    got = P.host_table(c, "trellis").reshape(-1, 2)
    assert np.array_equal(got[:, 0], ref[:, 0])
    m = ref[:, 0] != 65
    assert np.array_equal(got[m, 1], ref[m, 1])
    gexp = np.fromfile(os.path.join(GOLDEN, "rs_exp.u8"), dtype=np.uint8)
    assert np.array_equal(P.host_table(c, "rs_exp")[:511], gexp)


def test_qpsk_table_cells_follow_from_arithmetic(product):
    """Ground for a lighter critical path in k_rx: over ALL 65536 cells of the QPSK table the
    symbol is the pair of sign bits of the truncated (I, Q) and the cost is sat(d_nearest) - sat(d_second) with
    d_second = d_nearest + 212 min(abs I, abs Q) -- only phase_error (glibc atan2f) needs a table."""
    P = product
    cells = P.host_table(P.default_config(), "cstln").view(np.int16).reshape(256, 256, 4).astype(np.int32)
    v = np.arange(256).astype(np.uint8).view(np.int8).astype(np.int32)
    I, Q = v[:, None] * np.ones((1, 256), np.int32), v[None, :] * np.ones((256, 1), np.int32)
    assert np.array_equal(cells[:, :, 1], ((I < 0).astype(np.int32) << 1) | (Q < 0).astype(np.int32))
    d1 = (np.abs(I) - 53) ** 2 + (np.abs(Q) - 53) ** 2
    d2 = d1 + 212 * np.minimum(np.abs(I), np.abs(Q))
    assert np.array_equal(cells[:, :, 0], np.minimum(d1, 32767) - np.minimum(d2, 32767))


def test_first_batch_fir_taps_follow_tune(product, oracle):
    """fir_filter::set_freq at the frequency its first run() picks up from the demodulator (dsp.h:236-244, 270-280):
    the library's host code against the oracle, whose output with these taps equals the tapped reference
    (tests/test_oracle_cpu.py::test_oracle_resample_follows_tune_like_the_reference)."""
    P, O = product, oracle
    for Ftune, expect_shift in ((216000.0, True), (20000.0, False), (0.0, False)):
        got = P.host_table(P.default_config(fmt="f32", resample=True, Ftune=Ftune), "fir_shifted").view(np.float32).reshape(-1, 2)
        ch = O.Chain(O.Config(fmt="f32", resample=True, Ftune=Ftune))
        f = O.Fir(ch.fir_taps, ch.decim)
        if expect_shift:
            f.set_freq(float(np.float32(np.float32(Ftune) / np.float32(2.4e6))))
        want = f.shifted()
        assert got.shape == want.shape == (5, 2)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), Ftune
        assert bool(np.any(got[:, 1] != 0)) == expect_shift


@pytest.mark.parametrize("fec", ["1/2", "2/3", "4/6", "3/4", "5/6", "7/8"])
def test_viterbi_rescan_over_distinct_predecessors_selects_the_same_branch(product, fec):
    """k_viterbi.cu replaces the reference's rescan over ALL labels (viterbi.h:221-234, `<=`: the last label among
    equal metrics wins) by a scan over the largest label of every distinct predecessor, in increasing label order.
    Same winner (metric, predecessor, uncoded symbol) for every state, received label and metric vector -- including
    vectors full of ties -- on all six trellises."""
    P = product
    t = P.host_table(P.default_config(fec=fec), "trellis").reshape(64, -1, 2).astype(np.int32)
    ncs = t.shape[1]
    pred, us = t[:, :, 0], t[:, :, 1]
    # the compact lists, built like the kernel does: labels downwards, first sighting of a predecessor, reversed
    lists = []
    for s in range(64):
        seen, l = set(), []
        for c in range(ncs - 1, -1, -1):
            p = pred[s, c]
            if p == 65 or p in seen:
                continue
            seen.add(p)
            l.append((p, us[s, c]))
        lists.append(l[::-1])
    nb = len(lists[0])
    assert all(len(l) == nb for l in lists) and nb == min(64, 1 << int(fec[0]) if fec != "4/6" else 16)
    rng = np.random.default_rng(5)
    for trial in range(60):
        cc = rng.integers(0, 3 if trial % 2 else 1000, 64)          # every other trial: metrics full of ties
        bcost = -int(rng.integers(0, 3 if trial % 2 else 500))
        cs = int(rng.integers(0, ncs))
        for s in range(64):
            # reference order: the received label first (metric + cost), then every existing label ascending, `<=`
            best = (0x7fffffff, 0, 0)
            p = pred[s, cs]
            if p != 65 and cc[p] + bcost <= best[0]:
                best = (cc[p] + bcost, p, us[s, cs])
            full = best
            for c in range(ncs):
                p = pred[s, c]
                if p != 65 and cc[p] <= full[0]:
                    full = (cc[p], p, us[s, c])
            compact = best
            for p, u in lists[s]:
                if cc[p] <= compact[0]:
                    compact = (cc[p], p, u)
            assert tuple(int(v) for v in full) == tuple(int(v) for v in compact), (fec, s, cs, trial)


def test_viterbi_full_trellis_shortcut_selects_the_same_branch(product):
    """k_viterbi<kVitFull> (7/8): every state is a predecessor of every state, so the reference's rescan
    (viterbi.h:221-234) is not walked at all: its outcome is the minimum g of the current metrics, taken unless the
    labelled branch is STRICTLY smaller, and among the states that attain g the one whose entry comes last in this
    state's label order.  Same (metric, predecessor, uncoded symbol) as the reference's scan over all 256 labels for
    every state, received label and metric vector: unique minima, ties, the all-equal cold start, negative, zero and
    positive branch costs."""
    P = product
    t = P.host_table(P.default_config(fec="7/8"), "trellis").reshape(64, -1, 2).astype(np.int32)
    ncs = t.shape[1]
    pred, us = t[:, :, 0], t[:, :, 1]
    assert ncs == 256
    # the precondition the host checks before it selects the kernel (vit_trellis_is_full)
    for s in range(64):
        assert set(int(p) for p in pred[s] if p != 65) == set(range(64))
    # per state: position of predecessor p in the list "largest label of every distinct predecessor, ascending" and
    # the uncoded symbol of that entry -- the two transposed tables of the kernel
    pos = np.zeros((64, 64), np.int32)
    usp = np.zeros((64, 64), np.int32)
    for s in range(64):
        seen, k = set(), 64
        for c in range(ncs - 1, -1, -1):
            p = int(pred[s, c])
            if p == 65 or p in seen:
                continue
            seen.add(p)
            k -= 1
            pos[s, p], usp[s, p] = k, us[s, c]
        assert k == 0
    rng = np.random.default_rng(11)
    for trial in range(80):
        kind = trial % 4
        if kind == 0:
            cc = rng.integers(0, 1000, 64)                          # a unique minimum, almost surely
        elif kind == 1:
            cc = rng.integers(0, 3, 64)                             # many ties
        elif kind == 2:
            cc = np.zeros(64, np.int64)                             # cold start: all tied
        else:
            cc = rng.integers(0, 1000, 64); cc[rng.integers(0, 64, 3)] = 0   # a few tied minima (normalised metrics)
        bcost = int(rng.integers(-500, 1)) if trial % 5 else int(rng.integers(0, 3))
        cs = int(rng.integers(0, ncs))
        g = int(cc.min())
        tied = [p for p in range(64) if cc[p] == g]
        for s in range(64):
            ref = (0x7fffffff, 0, 0)                                # viterbi.h:207-234
            p = int(pred[s, cs])
            if p != 65 and cc[p] + bcost <= ref[0]:
                ref = (int(cc[p]) + bcost, p, int(us[s, cs]))
            for c in range(ncs):
                p = int(pred[s, c])
                if p != 65 and cc[p] <= ref[0]:
                    ref = (int(cc[p]), p, int(us[s, c]))
            p = int(pred[s, cs])                                    # the kernel
            if p != 65 and cc[p] + bcost < g:
                got = (int(cc[p]) + bcost, p, int(us[s, cs]))
            else:
                q = tied[0] if len(tied) == 1 else max(tied, key=lambda x: pos[s, x])
                got = (g, q, int(usp[s, q]))
            assert ref == got, (s, cs, trial, ref, got)
