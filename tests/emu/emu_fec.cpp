// emu_fec.cpp -- TEST INFRASTRUCTURE.  Runs the packet kernels of leansdr_b200/csrc/k_fec.cu on the host (cuda_emu.h):
// the Reed-Solomon decoder with and without the fused de-interleaver gather (k_rs), the byte re-alignment (k_realign),
// the sync flags (k_sync_flags) and the de-randomiser's output grid behind its scan (k_derand_*).  The device text is
// cut out of k_fec.cu by the test (everything inside its anonymous namespace; the launchers stay behind) and included
// here as FEC_DEV_INC.  The checker is the oracle's C restatement (liboracle.so): RS packets, flags and corrected-bit
// counts, de-randomised TS.  Built with -fsanitize=thread the same run is the race check of these kernels.
// `deconv`: the algebraic deconvolver (k_deconv_tiled + the carry thread k_deconv) against the oracle's deconvol_sync for
// every code rate and each of the four sync hypotheses, two consecutive batches (carried shift register, leftover
// bits, unread symbols), symbol bases at every 4-byte alignment.
// `sync`: the MPEG sync tracker (k_sync_flags + k_sync_track + k_realign driven pass by pass like run_sync of pipeline.cu)
// against the oracle's mpeg_sync on streams with a bit offset, either polarity, garbage in front (short: a search that
// locks; long: three fruitless sweeps, next_sync), a burst that loses the lock and a re-acquisition.
// `hs`: the hard-decision deconvolver of the --hs path (k_hs_errors + k_hs_lock + k_hs_decode) against the oracle's
// dvb_deconvol_sync_hard: votes every 32 chunks and every chunk, two batches (carried history, vote phase, alignment).
// Usage: emu_fec <case: rs|rs_deint|realign|derand|deconv|sync|hs> <seed>; exit code 0 = equal.
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../../leansdr_b200/csrc/kernels.h"
#include "../../leansdr_b200/csrc/tables.h"
extern "C" {
#include "../../oracle/dvbs_oracle.h"
}

namespace ldvb {
namespace dev {
#include FEC_DEV_INC
}
}  // namespace ldvb
using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

static void case_rs(uint64_t seed, bool deint) {
  std::mt19937_64 rng(seed);
  const size_t np = 64;
  uint8_t gexp[512], glog[256];
  make_rs_tables(gexp, glog);
  std::vector<uint8_t> code(204 * np), msg(188 * np);
  for (size_t p = 0; p < np; ++p) {
    for (int i = 0; i < 188; ++i) code[204 * p + i] = msg[188 * p + i] = (uint8_t)rng();
    orc_rs_encode(&code[204 * p]);
    const int ne = (int)(p % 11);          // 0..10 byte errors: up to 8 correctable, 9 and 10 not
    std::vector<int> pos;
    while ((int)pos.size() < ne) { const int q = (int)(rng() % 204); bool dup = false; for (int x : pos) dup |= x == q; if (!dup) pos.push_back(q); }
    for (int q : pos) code[204 * p + q] ^= (uint8_t)(1 + rng() % 255);
  }
  // reference: the oracle, packet by packet
  std::vector<uint8_t> want(188 * np);
  std::vector<int> want_bad(np), want_bits(np);
  for (size_t p = 0; p < np; ++p) {
    uint8_t pin[204];
    memcpy(pin, &code[204 * p], 204);
    int bits = 0;
    want_bad[p] = orc_rs_decode_packet(pin, &want[188 * p], &bits);
    want_bits[p] = bits;
  }
  // the kernel's input: plain packets, or the stream in front of the de-interleaver (dvb.h:936-941: byte i of packet p
  // sits at 204 * (p + i % 12) + i)
  std::vector<uint8_t> src;
  if (deint) {
    src.assign(204 * (np + 12), 0);
    for (size_t p = 0; p < np; ++p) for (int i = 0; i < 204; ++i) src[204 * (p + (size_t)(i % 12)) + i] = code[204 * p + i];
  } else {
    src = code;
  }
  std::vector<uint8_t> rs_out(204 * np, 0xee), rts(188 * np, 0xee);
  std::vector<int32_t> flags(2 * np, -1);
  const unsigned blocks = 5;               // (not a divisor of the packet count: the grid-stride loop wraps unevenly)
  if (deint) emu::launch(blocks, 128, [&] { dev::k_rs<true>(src.data(), np, gexp, glog, rs_out.data(), rts.data(), flags.data()); });
  else emu::launch(blocks, 128, [&] { dev::k_rs<false>(src.data(), np, gexp, glog, nullptr, rts.data(), flags.data()); });
  size_t accepted = 0;
  for (size_t p = 0; p < np; ++p) {
    CHECK((flags[2 * p] != 0) == (want_bad[p] != 0), "packet %zu: corrupted flag %d vs %d", p, flags[2 * p], want_bad[p]);
    if (want_bad[p]) continue;             // (the reference's output for uncorrectable packets is not defined: DESIGN section 2)
    ++accepted;
    CHECK(flags[2 * p + 1] == want_bits[p], "packet %zu: corrected bits %d vs %d", p, flags[2 * p + 1], want_bits[p]);
    CHECK(memcmp(&rts[188 * p], &want[188 * p], 188) == 0, "packet %zu: bytes differ from the oracle", p);
    CHECK(memcmp(&rts[188 * p], &msg[188 * p], 188) == 0, "packet %zu: bytes differ from the message", p);
  }
  CHECK(accepted >= np * 8 / 11, "only %zu packets accepted", accepted);
  if (deint) CHECK(memcmp(rs_out.data(), code.data(), 204 * np) == 0, "de-interleaved packets");
}

static void case_realign(uint64_t seed) {
  std::mt19937_64 rng(seed);
  const uint64_t np = 70, n = 204 * np;
  std::vector<uint8_t> bytes(n + 16);
  for (auto &b : bytes) b = (uint8_t)rng();
  for (int bitphase = 0; bitphase < 8; bitphase += 3) {
    for (int polarity : {0, 0xff}) {
      std::vector<uint8_t> out(n, 0xee);
      emu::launch((unsigned)((n + 255) / 256), 256, [&] { dev::k_realign(bytes.data(), n, bitphase, polarity, out.data()); });
      // mpeg_sync's byte at bit phase b (dvb.h:846-851): 16-bit window, shifted, polarity applied
      for (uint64_t i = 0; i < n; ++i) {
        const unsigned w = ((unsigned)bytes[i] << 8) | bytes[i + 1];
        const uint8_t want = (uint8_t)(((w >> bitphase) & 0xffu) ^ (unsigned)polarity);
        if (out[i] != want) { CHECK(false, "realign byte %llu phase %d polarity %d: %02x vs %02x", (unsigned long long)i, bitphase, polarity, out[i], want); break; }
      }
      SyncState st{};
      st.bitphase = bitphase; st.polarity = polarity; st.phase8 = (int)(rng() % 8);
      std::vector<uint32_t> words((np + 31) / 32 + 2, 0xeeeeeeeeu);
      emu::launch((unsigned)((np + 255) / 256), 256, [&] { dev::k_sync_flags(bytes.data(), np, &st, words.data()); });
      for (uint64_t p = 0; p < np; ++p) {
        const unsigned w = ((unsigned)bytes[204 * p] << 8) | bytes[204 * p + 1];
        const unsigned b = ((w >> bitphase) & 0xffu) ^ (unsigned)polarity;
        const unsigned expected = ((st.phase8 + p) & 7) ? 0x47u : 0xb8u;
        const bool bad = b != expected;
        CHECK((((words[p >> 5] >> (p & 31)) & 1) != 0) == bad, "sync flag of packet %llu", (unsigned long long)p);
      }
    }
  }
}

static void case_derand(uint64_t seed) {
  // RS-decoded packets of a real randomised stream (the oracle's randomiser), some of them damaged or marked, through
  // the scan grids and the XOR grid; the checker is the oracle's derandomizer
  std::mt19937_64 rng(seed);
  const uint64_t np = 700;
  const std::vector<uint8_t> pattern = make_derand_pattern();
  std::vector<uint8_t> rts(188 * np);
  for (uint64_t p = 0; p < np; ++p) {
    rts[188 * p] = 0x47;
    for (int i = 1; i < 188; ++i) rts[188 * p + i] = (uint8_t)rng();
    const int ph = (int)(p % 8);
    for (int i = 0; i < 188; ++i) rts[188 * p + i] ^= pattern[188 * ph + i];   // randomise: sync of packet 0 of 8 becomes 0xb8
    const uint64_t r = rng() % 40;
    if (r == 0) rts[188 * p] ^= 0x55;            // what rs_decoder does to a packet it could not correct
    else if (r == 1) rts[188 * p] = (uint8_t)rng();
  }
  orc_derand od; od.pos = 188 * (int)(rng() % 8);
  std::vector<uint8_t> want(188 * np);
  const int pos_in = od.pos;
  const size_t nwant = orc_derandomize(&od, rts.data(), np, want.data());
  std::vector<uint32_t> scratch(2 * np + 16 * (np / 64 + 1) + 64, 0);
  std::vector<uint64_t> counts(4, 99);
  std::vector<uint8_t> ts(188 * np, 0xee);
  DerandArgs a{};
  a.rts = rts.data(); a.npackets = np; a.pattern = pattern.data(); a.pos_in = pos_in; a.ts_out = ts.data(); a.ts_cap = np;
  a.counts = counts.data(); a.flags = nullptr; a.scratch = scratch.data();
  const unsigned tile = 64, ntiles = (unsigned)((np + tile - 1) / tile);
  emu::launch(ntiles, tile, [&] { dev::k_derand_tiles(a); });
  emu::launch(1, 32, [&] { dev::k_derand_chain(a, ntiles, tile); });
  emu::launch(ntiles, tile, [&] { dev::k_derand_index(a); });
  emu::launch((unsigned)((np + 3) / 4), 256, [&] { dev::k_derand_out(a); });
  CHECK(counts[0] == nwant, "kept %llu vs %zu", (unsigned long long)counts[0], nwant);
  CHECK((int)counts[2] == od.pos, "carried position %llu vs %d", (unsigned long long)counts[2], od.pos);
  CHECK(nwant > np / 2, "degenerate stream: %zu kept", nwant);
  if (counts[0] == nwant) CHECK(memcmp(ts.data(), want.data(), 188 * nwant) == 0, "TS bytes");
}

static void case_deconv(uint64_t seed) {
  std::mt19937_64 rng(seed);
  for (int fec = 0; fec <= 5; ++fec) {
    DeconvPolys dp;
    if (!make_deconv(fec, &dp)) continue;               // (4/6 is a Viterbi-only trellis)
    for (int locked = 0; locked < 4; ++locked) {
      orc_deconv od;
      orc_deconv_init(&od, fec);
      orc_deconv_set(&od, locked, 0);
      CHECK(od.punctperiod == dp.punctperiod && od.punctweight == dp.punctweight, "puncturing of fec %d", fec);
      uint64_t reg = 0, acc = 0; int n_in = 0, n_out = 0;          // the product's carried state (HypState of pipeline.cu)
      std::vector<uint32_t> unread;                                // symbols the previous batch left behind
      for (int batch = 0; batch < 2; ++batch) {
        const int mis = (int)(rng() % 4);                          // alignment of the stream's read position
        const size_t fresh = batch ? 2000 + rng() % 3000 : 17000 + rng() % 9000;   // (first batch: several 1024-byte CTAs at every rate)
        std::vector<uint32_t> buf(mis + unread.size() + fresh + 8, 0xdeadbeefu);
        uint32_t *sym = buf.data() + mis;
        for (size_t i = 0; i < unread.size(); ++i) sym[i] = unread[i];
        for (size_t i = 0; i < fresh; ++i) sym[unread.size() + i] = (uint32_t)(rng() % 4) << 16 | (uint32_t)(rng() & 0xffff);
        const size_t count = unread.size() + fresh;
        const uint64_t nbytes = (count - 64) / (dp.punctweight / 2) * dp.punctperiod / 8;   // dvb.h:420
        DeconvArgs a{};
        a.symbols = sym; a.nbytes = nbytes; a.reg_in = reg; a.n_in = n_in; a.out_acc = acc; a.n_out = n_out;
        for (int k = 0; k < 4; ++k) a.hyp[k] = dp.hyp_lut[locked][k];
        a.punctperiod = dp.punctperiod; a.punctweight = dp.punctweight;
        for (int b = 0; b < 8; ++b) a.deconv[b] = dp.deconv[b];
        std::vector<uint8_t> out(nbytes + 8, 0xee);
        a.out = out.data();
        uint64_t carry[5] = {0, 0, 0, 0, 0};
        emu::launch((unsigned)((nbytes + dev::kDcBytes - 1) / dev::kDcBytes), 256, [&] { dev::k_deconv_tiled(a, count); });
        emu::launch(1, 32, [&] { dev::k_deconv(a, carry); });
        std::vector<uint8_t> want(nbytes + 8, 0xee);
        size_t consumed = 0;
        const size_t nw = orc_deconv_run2(&od, reinterpret_cast<const uint8_t *>(sym), count, want.data(), nbytes, &consumed, 1);
        CHECK(nw == nbytes, "fec %d hyp %d batch %d: oracle wrote %zu of %llu bytes", fec, locked, batch, nw, (unsigned long long)nbytes);
        CHECK(memcmp(out.data(), want.data(), nbytes) == 0, "fec %d hyp %d batch %d: bytes differ from the oracle", fec, locked, batch);
        CHECK(out[nbytes] == 0xee, "fec %d hyp %d batch %d: wrote past the end", fec, locked, batch);
        CHECK(carry[4] == consumed, "fec %d hyp %d batch %d: consumed %llu vs oracle %zu", fec, locked, batch, (unsigned long long)carry[4], consumed);
        const orc_dsync &os = od.syncs[locked];
        CHECK((int)(int64_t)carry[1] == os.n_in && (int)(int64_t)carry[3] == os.n_out, "fec %d hyp %d batch %d: register fill %d/%d vs oracle %d/%d",
              fec, locked, batch, (int)(int64_t)carry[1], (int)(int64_t)carry[3], os.n_in, os.n_out);
        const uint64_t mo = os.n_out ? ((os.n_out >= 64) ? ~0ull : ((1ull << os.n_out) - 1)) : 0;
        CHECK((carry[2] & mo) == (os.out & mo), "fec %d hyp %d batch %d: leftover bits", fec, locked, batch);
        CHECK(carry[0] == os.in, "fec %d hyp %d batch %d: shift register %016llx vs oracle %016llx", fec, locked, batch,
              (unsigned long long)carry[0], (unsigned long long)os.in);
        reg = carry[0]; n_in = (int)(int64_t)carry[1]; acc = carry[2]; n_out = (int)(int64_t)carry[3];
        unread.assign(sym + carry[4], sym + count);
      }
    }
  }
}

static void case_sync(uint64_t seed) {
  std::mt19937_64 rng(seed);
  for (int variant = 0; variant < 4; ++variant) {
    // the byte stream mpeg_sync reads: packets with 0x47 / 0xb8 heads, inverted or not, behind a bit offset and garbage
    const int bitoff = (int)(rng() % 8);
    const bool inverted = (variant & 1) != 0;
    const size_t garbage = (variant >= 2) ? 41000 + rng() % 3000 : rng() % 3000;   // (3 sweeps of 8 phases = 39 168 bytes)
    const size_t npk = 300;   // (a search walks the 8 bit phases one 8-packet window at a time: up to 64 packets to lock)
    std::vector<uint8_t> pk(204 * npk);
    const size_t phase0 = rng() % 8;
    for (size_t p = 0; p < npk; ++p) {
      for (int i = 0; i < 204; ++i) pk[204 * p + i] = (uint8_t)rng();
      pk[204 * p] = ((p + phase0) % 8 == 0) ? 0xb8 : 0x47;
      if (p >= 150 && p < 156) pk[204 * p] ^= 0x21;             // a burst of bad syncs: the lock times out, then a new search
      if (p == 100 || p == 260) pk[204 * p] ^= 0x04;            // isolated bad syncs: the lock holds
    }
    std::vector<uint8_t> in(garbage + pk.size() + 2);
    for (size_t i = 0; i < garbage; ++i) in[i] = (uint8_t)rng();
    {  // packets shifted by bitoff bits: byte k of the stream = (window of pk bytes k-1, k) >> (8 - bitoff) style
      uint32_t w = (uint8_t)rng();
      for (size_t k = 0; k < pk.size() + 2; ++k) {
        const uint8_t b = k < pk.size() ? pk[k] : (uint8_t)rng();
        w = (w << 8) | b;
        in[garbage + k] = (uint8_t)(w >> bitoff);
      }
    }
    if (inverted) for (size_t k = garbage; k < in.size(); ++k) in[k] ^= 0xff;
    // product state (ldvb_create / reset_carry) and oracle state
    SyncState st{};
    st.report_state = 1; st.phase8 = -1; st.fastlock = 0; st.resync_period = 1;
    orc_mpegsync om;
    orc_mpegsync_init(&om);
    std::vector<uint8_t> got, want;
    size_t pos_p = 0, pos_o = 0;
    for (int pass = 0; pass < 200; ++pass) {
      // one pass of the product over the unread bytes (run_sync)
      const uint8_t *bytes = in.data() + pos_p;
      const uint64_t count = in.size() - pos_p;
      uint64_t nflag = 0;
      bool product_idle = false;
      std::vector<uint32_t> words((count / 204) / 32 + 4, 0);
      SyncResult r{};
      if (st.synchronized) {
        nflag = count >= 205 ? (count - 1) / 204 : 0;
        if (!nflag) product_idle = true;
        else emu::launch((unsigned)((nflag + 255) / 256), 256, [&] { dev::k_sync_flags(bytes, nflag, &st, words.data()); });
      } else if (count < 204 * 8 + 1) {
        product_idle = true;
      }
      if (!product_idle) {
        emu::launch(1, 256, [&] { dev::k_sync_track(bytes, count, &st, words.data(), nflag, &r); });
        if (r.produced) {
          const size_t at = got.size();
          got.resize(at + r.produced);
          // (k_realign is one host thread per byte here and is checked byte for byte in case_realign: in long runs
          //  only the two ends go through the kernel, the middle through mpeg_sync's formula)
          const uint64_t edge = 1024;
          if (r.produced <= 2 * edge) {
            emu::launch((unsigned)((r.produced + 255) / 256), 256, [&] { dev::k_realign(bytes, r.produced, st.bitphase, st.polarity, got.data() + at); });
          } else {
            emu::launch((unsigned)(edge / 256), 256, [&] { dev::k_realign(bytes, edge, st.bitphase, st.polarity, got.data() + at); });
            for (uint64_t i = edge; i < r.produced - edge; ++i)
              got[at + i] = (uint8_t)((((((unsigned)bytes[i] << 8) | bytes[i + 1]) >> st.bitphase) & 0xffu) ^ (unsigned)(st.polarity & 0xff));
            const uint64_t tail0 = r.produced - edge;
            emu::launch((unsigned)(edge / 256), 256, [&] { dev::k_realign(bytes + tail0, edge, st.bitphase, st.polarity, got.data() + at + tail0); });
          }
        }
        st = r.st;
        pos_p += r.consumed;
      }
      // one run() of the oracle over ITS unread bytes
      std::vector<uint8_t> o(in.size() + 256);
      size_t consumed = 0, nl = 0, nlt = 0; int lock_ev[64]; int switched = 0;
      std::vector<uint64_t> lt(in.size() / 204 + 8);
      const size_t nw = orc_mpegsync_run2(&om, nullptr, in.data() + pos_o, in.size() - pos_o, o.data(), o.size(), &consumed, lock_ev, &nl,
                                          lt.data(), &nlt, 1, &switched);
      want.insert(want.end(), o.begin(), o.begin() + nw);
      pos_o += consumed;
      CHECK(pos_p == pos_o, "variant %d pass %d: consumed %zu vs oracle %zu", variant, pass, pos_p, pos_o);
      CHECK(got.size() == want.size(), "variant %d pass %d: produced %zu vs oracle %zu", variant, pass, got.size(), want.size());
      CHECK((!product_idle && r.need_next_sync) == (switched != 0), "variant %d pass %d: next_sync %d vs oracle %d", variant, pass, r.need_next_sync, switched);
      CHECK(st.synchronized == om.synchronized && st.bitphase == om.bitphase && st.next_sync_count == om.next_sync_count, "variant %d pass %d: state", variant, pass);
      if (st.synchronized) CHECK((st.polarity & 0xff) == om.polarity && st.phase8 == om.phase8 && st.lock_timeleft == om.lock_timeleft && st.locktime == om.locktime,
                                 "variant %d pass %d: lock state (phase8 %d/%d, left %u/%lu, time %llu/%lu)", variant, pass, st.phase8, om.phase8,
                                 st.lock_timeleft, om.lock_timeleft, (unsigned long long)st.locktime, om.locktime);
      if (g_fail) return;
      if ((product_idle || (!r.consumed && !r.produced)) && !consumed && !nw) break;
    }
    CHECK(got == want, "variant %d: aligned bytes", variant);
    CHECK(got.size() >= 204 * 120, "variant %d: only %zu aligned bytes", variant, got.size());
    // what came out are the packets themselves (from the first one the tracker locked on)
    bool found = false;
    for (size_t p = 0; p + 1 < npk && !found; ++p) found = got.size() >= 204 && memcmp(got.data(), pk.data() + 204 * p, 204) == 0;
    CHECK(found, "variant %d: the aligned stream does not start with a transmitted packet", variant);
  }
}

static void case_hs(uint64_t seed) {
  std::mt19937_64 rng(seed);
  for (int period : {32, 1}) {
    orc_hsdeconv od;
    orc_hsdeconv_init(&od, period);
    uint64_t hist = 0; int hist_valid = 0, phase = 0, locked = 0;     // the product's carry (pipeline.cu, --hs branch)
    for (int batch = 0; batch < 2; ++batch) {
      const uint64_t nchunks = batch ? 37 + rng() % 20 : 70 + rng() % 30;
      const uint64_t nsym = nchunks * 512;
      std::vector<uint32_t> words(nsym);
      std::vector<uint8_t> syms(nsym);
      // runs of one symbol sequence repeated make some alignments clearly better than others; noise in between
      for (uint64_t i = 0; i < nsym; ++i) { syms[i] = (uint8_t)((i / 3000) % 2 ? rng() % 4 : (i * 7 + i / 5) % 4); words[i] = (uint32_t)syms[i] << 16 | (uint32_t)(rng() & 0xffff); }
      const uint64_t ngroups_max = nchunks / period + 2;
      std::vector<uint32_t> errors(ngroups_max * 4 + 16, 0);
      std::vector<uint8_t> lock_of_chunk(nchunks + 16, 0xee), out(nchunks * 64 + 8, 0xee);
      int32_t state_out[4] = {-1, -1, -1, -1};
      HsDeconvArgs a{};
      a.symbols = words.data(); a.nchunks = nchunks; a.hist = hist; a.hist_valid = hist_valid;
      a.resync_phase = phase; a.resync_period = period; a.locked = locked;
      a.errors = errors.data(); a.lock_of_chunk = lock_of_chunk.data(); a.out = out.data(); a.state_out = state_out;
      // launch_hs_deconv (k_fec.cu)
      const uint64_t first = (uint64_t)((period - phase) % period);
      const uint32_t ngroups = first < nchunks ? (uint32_t)((nchunks - first + period - 1) / period) : 0;
      if (ngroups) emu::launch(ngroups, 128, [&] { dev::k_hs_errors(a, first, ngroups); });
      emu::launch(1, 32, [&] { dev::k_hs_lock(a, ngroups); });
      emu::launch((unsigned)((nchunks * 64 + 255) / 256), 256, [&] { dev::k_hs_decode(a, first); });
      std::vector<uint8_t> want(nchunks * 64 + 8, 0xee);
      size_t consumed = 0;
      const size_t nw = orc_hsdeconv_run(&od, syms.data(), nsym, want.data(), nchunks * 64, &consumed);
      CHECK(nw == nchunks * 64 && consumed == nsym, "period %d batch %d: oracle wrote %zu bytes from %zu symbols", period, batch, nw, consumed);
      CHECK(memcmp(out.data(), want.data(), nchunks * 64) == 0, "period %d batch %d: bytes differ from the oracle", period, batch);
      for (uint64_t i = 0; i < nchunks * 64 && g_fail < 4; ++i) CHECK(out[i] == want[i], "byte %llu (chunk %llu): %02x vs %02x", (unsigned long long)i, (unsigned long long)(i / 64), out[i], want[i]);
      CHECK(out[nchunks * 64] == 0xee, "period %d batch %d: wrote past the end", period, batch);
      CHECK(state_out[0] == od.locked, "period %d batch %d: alignment %d vs oracle %d", period, batch, state_out[0], od.locked);
      locked = state_out[0];
      phase = (int)((phase + nchunks) % (uint64_t)period);
      CHECK(phase == od.resync_phase, "period %d batch %d: vote phase", period, batch);
      hist = 0;
      for (int i = 0; i < 32; ++i) hist = (hist << 2) | ((words[nsym - 32 + i] >> 16) & 3u);
      hist_valid = 32;
    }
  }
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: emu_fec <rs|rs_deint|realign|derand|deconv|sync|hs> <seed>\n"); return 2; }
  const std::string c = argv[1];
  const uint64_t seed = strtoull(argv[2], nullptr, 10);
  if (c == "rs") case_rs(seed, false);
  else if (c == "rs_deint") case_rs(seed, true);
  else if (c == "realign") case_realign(seed);
  else if (c == "derand") case_derand(seed);
  else if (c == "deconv") case_deconv(seed);
  else if (c == "sync") case_sync(seed);
  else if (c == "hs") case_hs(seed);
  else { fprintf(stderr, "unknown case\n"); return 2; }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_fec %s seed %llu: equal\n", c.c_str(), (unsigned long long)seed);
  return 0;
}
