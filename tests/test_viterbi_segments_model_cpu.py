"""Model check (CPU) of the time-segment schedule of the CUDA Viterbi stage (leansdr_b200/csrc/k_viterbi.cu).

The schedule -- cut the stream at re-sync chunks, start every segment's decoders cold a little earlier, compare the
state with which a segment enters its own chunks with the predecessor's exit state bit for bit, re-run what did not
merge -- is replayed here with the oracle's block update (viterbi_dec::update restated, pinned to the reference by
tests/test_oracle_cpu.py), one chunk at a time:

  * exactness: stitching the segments with the verify/repair rule reproduces the serial decoder's bytes;
  * the measurement behind the warm-up lengths: how often a cold segment's entry state equals the true state, as a
    function of the number of re-sync chunks the non-current decoders warm up on (DESIGN.md "Viterbi").
"""
import ctypes as C

import numpy as np
import pytest

from tests import vectors as V

needs_ref = pytest.mark.skipif(not V.have_ref(), reason="oracle/_ref binaries not built")


class Model:
    def __init__(self, O, cst, fec, symbols, period=32):
        self.O, self.L = O, O.lib()
        L = self.L
        vp = C.c_void_p
        L.orc_viterbi_chunk.restype = C.c_size_t
        L.orc_viterbi_chunk.argtypes = [vp, vp, C.c_uint32, C.c_int, vp, vp]
        L.orc_viterbi_get_dec.argtypes = [vp, C.c_int, vp, vp]
        L.orc_viterbi_set_dec.argtypes = [vp, C.c_int, vp, vp]
        L.orc_viterbi_set_ctl.argtypes = [vp, C.c_int, C.c_int]
        for f in ("orc_viterbi_resync_phase", "orc_viterbi_nshifts", "orc_viterbi_bits_in"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_int
        self.cells, self.syms, self.cst = O.cstln_table(cst, False, fec)
        self.fec = "4/6" if fec == "2/3" and self.syms.shape[0] in (4, 64) else fec
        self.v = O.Viterbi(self.cst, self.fec)
        self.v.set_resync_period(period)
        self.P = period
        self.ns = self.v.nsyncs()
        self.nsh = L.orc_viterbi_nshifts(self.v.h)
        self.bits_in = L.orc_viterbi_bits_in(self.v.h)
        self.sym = np.ascontiguousarray(symbols, np.uint8).reshape(-1, 4)
        self.nchunks = (self.sym.shape[0] - (self.nsh - 1)) // (128 * self.nsh)
        self.bpc = 16 * self.bits_in
        self.all = (1 << self.ns) - 1

    def fresh(self):
        v = self.O.Viterbi(self.cst, self.fec)
        v.set_resync_period(self.P)
        return v

    def chunk(self, v, c, mask, out_sync):
        out = np.zeros(self.bpc, np.uint8)
        tot = np.zeros(self.ns, np.int32)
        p = self.sym[c * 128 * self.nsh:]
        n = self.L.orc_viterbi_chunk(v.h, p.ctypes.data_as(C.c_void_p), mask, out_sync,
                                     out.ctypes.data_as(C.c_void_p), tot.ctypes.data_as(C.c_void_p))
        return out[:n], tot

    def state(self, v):
        cost = np.zeros((self.ns, 64), np.int32)
        path = np.zeros((self.ns, 64), np.uint64)
        for s in range(self.ns):
            self.L.orc_viterbi_get_dec(v.h, s, cost[s].ctypes.data_as(C.c_void_p), path[s].ctypes.data_as(C.c_void_p))
        return cost, path

    def load(self, v, st):
        cost, path = st
        for s in range(self.ns):
            self.L.orc_viterbi_set_dec(v.h, s, np.ascontiguousarray(cost[s]).ctypes.data_as(C.c_void_p),
                                       np.ascontiguousarray(path[s]).ctypes.data_as(C.c_void_p))

    @staticmethod
    def vote(current, tot):
        best = current
        for s in range(len(tot)):
            if tot[s] > tot[best]:
                best = s
        return best

    def main_chunks(self, v, current, c0, c1, collect):
        """The segment's own chunks, as the kernel's M steps (phase0 = 0: chunk c is a re-sync chunk iff c % P == 0)."""
        out = []
        for c in range(c0, c1):
            resync = c % self.P == 0
            mask = self.all if resync else (1 << current)
            b, tot = self.chunk(v, c, mask, current)
            if collect:
                out.append(b)
            if resync:
                current = self.vote(current, tot)
        return current, out

    def serial(self, bounds):
        """One pass from the constructor state; snapshots (state, current) at the given chunk indices."""
        v = self.fresh()
        snaps, cur, outs, prev = {}, 0, [], 0
        for b in list(bounds) + [self.nchunks]:
            cur, o = self.main_chunks(v, cur, prev, b, True)
            outs += o
            snaps[b] = (self.state(v), cur)
            prev = b
        return np.concatenate(outs), snaps

    def cold_entry(self, c0, carried, warm_others, warm_chunks):
        """State with which a cold segment enters chunk c0 (the kernel's A and B steps)."""
        v = self.fresh()
        cur = carried[1]
        nA = c0 // self.P                       # re-sync chunks in [0, c0)
        if nA < warm_others:
            self.load(v, carried[0])            # close to the start: the other decoders follow exactly from the carried state
        else:
            nA = warm_others
        for j in range(nA, 0, -1):
            _, tot = self.chunk(v, c0 - j * self.P, self.all, -1)
            cur = self.vote(cur, tot)
        zero = (np.zeros(64, np.int32), np.zeros(64, np.uint64))
        self.L.orc_viterbi_set_dec(v.h, cur, zero[0].ctypes.data_as(C.c_void_p), zero[1].ctypes.data_as(C.c_void_p))
        for c in range(c0 - warm_chunks, c0):
            self.chunk(v, c, 1 << cur, -1)
        return v, cur


def _same(a, b):
    return a[1] == b[1] and np.array_equal(a[0][0], b[0][0]) and np.array_equal(a[0][1], b[0][1])


CASES = [
    ("qpsk12-noise", "QPSK", "1/2", dict(fmt="f32", viterbi=True), dict(noise_db=25), 500),
    ("qpsk78", "QPSK", "7/8", dict(fmt="f32", viterbi=True, fec="7/8", Fs=4e6), dict(cr="7/8", ratio="2"), 560),
    ("8psk23", "8PSK", "2/3", dict(fmt="f32", viterbi=True, cstln="8PSK", fec="2/3", Fs=4e6), dict(cr="2/3", ratio="2", cst="8PSK"), 500),
]


@needs_ref
@pytest.mark.parametrize("name,cst,fec,kw,gkw,npk", CASES, ids=[c[0] for c in CASES])
def test_segment_schedule_is_exact_and_merges(oracle, name, cst, fec, kw, gkw, npk):
    O = oracle
    raw = V.ref_iq(npk, fmt="f32", **gkw)
    ref = O.Chain(O.Config(**kw)).run(raw)
    m = Model(O, cst, fec, ref["symbols"])
    L = 64                                                 # chunks per segment (two re-sync groups)
    bounds = list(range(L, m.nchunks - L // 2, L))
    assert len(bounds) >= 4
    serial_bytes, snaps = m.serial(bounds)
    assert np.array_equal(serial_bytes, ref["bytes"][:serial_bytes.size])          # the model's serial pass IS the oracle
    carried = ((np.zeros((m.ns, 64), np.int32), np.zeros((m.ns, 64), np.uint64)), 0)   # constructor state at chunk 0

    # ---- the schedule of k_viterbi.cu: 16 re-sync chunks for every decoder, 2 chunks for the current one
    out = [m.main_chunks(m.fresh(), 0, 0, bounds[0], True)[1]]
    prev_exit, repaired = snaps[bounds[0]], 0
    ends = bounds[1:] + [m.nchunks]
    for c0, c1 in zip(bounds, ends):
        v, cur = m.cold_entry(c0, carried, 16, 2)
        entry = (m.state(v), cur)
        if not _same(entry, prev_exit):                    # k_vit_verify failed: re-run exactly from the predecessor's exit
            repaired += 1
            v = m.fresh(); m.load(v, prev_exit[0]); cur = prev_exit[1]
        cur, o = m.main_chunks(v, cur, c0, c1, True)
        out.append(o)
        prev_exit = (m.state(v), cur)
        assert _same(prev_exit, snaps[c1])                 # induction: exit states equal the serial pass
    stitched = np.concatenate([b for seg in out for b in seg])
    assert np.array_equal(stitched, serial_bytes)
    # the first segments sit in the acquisition transient (the current hypothesis is still being chosen)
    assert repaired <= max(2, len(bounds) // 4), (repaired, len(bounds))


@needs_ref
def test_wrong_hypothesis_decoders_need_a_long_warmup(oracle):
    """The measurement behind warm_others = 16: fraction of cold segment entries that equal the true state, by the
    number of re-sync chunks the decoders warm up on (QPSK 1/2: 1 right hypothesis, 3 that decode noise)."""
    O = oracle
    raw = V.ref_iq(2000, fmt="f32", noise_db=25)
    ref = O.Chain(O.Config(fmt="f32", viterbi=True)).run(raw)
    m = Model(O, "QPSK", "1/2", ref["symbols"])
    bounds = list(range(1024, m.nchunks - 32, 96))[:40]
    assert len(bounds) >= 30
    _, snaps = m.serial(bounds)
    carried = ((np.zeros((m.ns, 64), np.int32), np.zeros((m.ns, 64), np.uint64)), 0)
    frac = {}
    for k in (1, 4, 16):
        ok = 0
        for c0 in bounds:
            v, cur = m.cold_entry(c0, carried, k, 2)
            ok += _same((m.state(v), cur), snaps[c0])
        frac[k] = ok / len(bounds)
    print("cold entries equal to the true state, by re-sync chunks of warm-up:", frac)
    assert frac[16] >= 0.9 and frac[16] >= frac[1]
    assert frac[1] <= 0.5          # one re-sync chunk (128 blocks) is not enough for the noise-fed decoders
