// emu_ctl.cpp -- TEST INFRASTRUCTURE.  Runs the control kernels of leansdr_b200/csrc/k_ctl_*.cuh on the host (cuda_emu.h)
// and compares every output, bit for bit, with the kernels they replace (ctl_v1.cuh: the text that ran on B200 under the
// 144 GPU parity tests of commit f37a100) on seeded random inputs that cover the edge cases of each (empty and ragged
// batches, tile boundaries, carried state, resets, lock loss).  Usage: emu_ctl <case> [seed]; exit code 0 = identical.
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

#include "../../leansdr_b200/csrc/kernels.h"

namespace ldvb {
namespace v1 {
#include "ctl_v1.cuh"
}
namespace v2 {
__device__ __forceinline__ unsigned par64(uint64_t v) { return __popcll(v) & 1; }
#include "../../leansdr_b200/csrc/k_ctl_fec.cuh"
#include "../../leansdr_b200/csrc/k_ctl_rx.cuh"
}  // namespace v2
}  // namespace ldvb

using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

// ------------------------------------------------------------------------------------------------ k_rx_plan
static void case_plan(uint64_t seed) {
  std::mt19937_64 rng(seed);
  const uint32_t sizes[] = {1, 2, 5, 1023, 1024, 1025, 2047, 4100, 20011};
  for (uint32_t nspans : sizes) {
    for (int variant = 0; variant < 2; ++variant) {
      const int nrot = variant ? 8 : 4;
      std::vector<RxSpanInfo> info(nspans);
      std::vector<RxSeam> seams(nspans);
      const uint32_t span_cap = 140;
      for (auto &i : info) { i.n_out = 50 + rng() % 70; i.n_tail = rng() % 12; i.n_head_logged = rng() % 100; i.pad = 0; if (rng() % 997 == 0) i.n_tail = 200; }
      for (auto &s : seams) {
        s.ok = (rng() % 50) != 0 || variant == 0; s.rot = (int)(rng() % nrot); s.extend_prev = rng() & 1; s.skip_next = rng() & 1;
        s.compared = 100; s.mismatches = (rng() % 20 == 0) ? (int)(rng() % 5) : 0; s.ok_loose = (rng() % 80) != 0;
        s.dphase = (float)((int)(rng() % 2001) - 1000) * 0.37f; s.dfreqw = (float)((int)(rng() % 2001) - 1000) * 0.011f; s.dmu = (float)((int)(rng() % 2001) - 1000) * 1e-4f;
      }
      const int rot0 = variant ? (int)(rng() % nrot) : 0; const uint32_t skip0 = variant ? (uint32_t)(rng() & 1) : 0;
      std::vector<uint64_t> off1(nspans + 1, 77), off2(nspans + 1, 77), res1(9, 5), res2(9, 5);
      std::vector<uint32_t> sk1(nspans, 9), sk2(nspans, 9);
      std::vector<uint8_t> ro1(nspans, 9), ro2(nspans, 9);
      emu::launch(1, 1024, [&] { v1::k_rx_plan(info.data(), seams.data(), nspans, span_cap, nrot, rot0, skip0, off1.data(), sk1.data(), ro1.data(), res1.data()); });
      {   // the two grids of the current text (launch_rx_plan in k_rx.cu: result zeroed, totals behind the offsets);
          // 64 spans per CTA here, 1024 in the library: the same code, a fraction of the host threads
        const unsigned bs = 64, nblk = std::max(1u, (nspans + bs - 1) / bs);
        std::vector<unsigned long long> totals(2 * nblk, 99), r2(9, 0);
        emu::launch(nblk, bs, [&] { v2::k_rx_plan_local(info.data(), seams.data(), nspans, span_cap, nrot, rot0, skip0, off2.data(), sk2.data(), ro2.data(), totals.data(), r2.data()); });
        emu::launch(nblk, bs, [&] { v2::k_rx_plan_apply(nspans, nrot, off2.data(), ro2.data(), totals.data(), r2.data()); });
        for (int k = 0; k < 9; ++k) res2[k] = r2[k];
      }
      CHECK(off1 == off2, "plan offsets nspans=%u", nspans);
      CHECK(sk1 == sk2, "plan skips nspans=%u", nspans);
      CHECK(ro1 == ro2, "plan rots nspans=%u", nspans);
      for (int k = 0; k < 9; ++k) CHECK(res1[k] == res2[k], "plan result[%d] nspans=%u: %llu vs %llu", k, nspans, (unsigned long long)res1[k], (unsigned long long)res2[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ k_derand_scan
static void case_derand(uint64_t seed) {
  std::mt19937_64 rng(seed);
  const uint64_t sizes[] = {1, 7, 63, 64, 65, 1000, 1023, 1024, 1025, 4097};   // (4097: 65 tiles of 64, three rounds of the 32-thread chain below)
  for (uint64_t np : sizes) {
    for (int variant = 0; variant < 4; ++variant) {
      if (variant == 3 && np < 1000) continue;
      std::vector<uint8_t> pattern(1504);
      for (auto &b : pattern) b = (uint8_t)rng();
      if (variant != 1) for (int k = 0; k < 8; ++k) pattern[188 * k] = 0;   // (sync bytes are not randomised on air)
      std::vector<uint8_t> rts(188 * np);
      // a plausible stream: resets every 8 packets from a random origin (none at all in variant 2 for the small sizes), damage here and there
      const uint64_t origin = rng() % 8;
      for (uint64_t p = 0; p < np; ++p) {
        uint8_t head = ((p + 8 - origin) % 8 == 0) ? 0xb8 : 0x47;
        if (variant == 2 && np < 5000) head = 0x47;
        const uint64_t r = rng() % 64;
        if (r == 0) head = (uint8_t)rng();
        else if (r == 1) head = 0xb8 ^ 0x55;
        else if (r == 2) head = 0xb8;
        else if (r == 3) head ^= 0x55;
        // variant 3: whole tiles without a reset (none after packet 300; none at all for one of the sizes):
        // the pattern position of every later tile follows from a reset several tiles back, or from the carried one
        if (variant == 3 && (p >= 300 || np == 1025) && (head == 0xb8 || head == (0xb8 ^ 0x55))) head = 0x47;
        rts[188 * p] = head;
      }
      std::vector<int32_t> flags(2 * np);
      for (auto &f : flags) f = (int32_t)(rng() % 9);
      DerandArgs a{};
      a.rts = rts.data(); a.npackets = np; a.pattern = pattern.data(); a.pos_in = 188 * (int32_t)(rng() % 8);
      a.ts_out = nullptr; a.ts_cap = np; a.flags = (variant == 1) ? nullptr : flags.data();
      std::vector<uint32_t> s1(2 * np + 16 * (np / 64 + 1) + 64, 0xabababab), s2(s1);   // (+ the tile records of the current text)
      std::vector<uint64_t> c1(4, 99), c2(4, 99);
      DerandArgs a1 = a, a2 = a;
      a1.scratch = s1.data(); a1.counts = c1.data(); a2.scratch = s2.data(); a2.counts = c2.data();
      emu::launch(1, 1024, [&] { v1::k_derand_scan(a1); });
      {   // the three grids of the current text (launch_derand in k_fec.cu); tiles of 64 packets here, 1024 in the
          // library, and 32 tiles per round of the chain: the same code, a fraction of the host threads
        const unsigned tile = 64, ntiles = (unsigned)((np + tile - 1) / tile);
        emu::launch(ntiles, tile, [&] { v2::k_derand_tiles(a2); });
        emu::launch(1, 32, [&] { v2::k_derand_chain(a2, ntiles, tile); });
        emu::launch(ntiles, tile, [&] { v2::k_derand_index(a2); });
      }
      CHECK(std::equal(s1.begin(), s1.begin() + 2 * np, s2.begin()), "derand scratch np=%llu variant=%d", (unsigned long long)np, variant);
      CHECK(c1 == c2, "derand counts np=%llu variant=%d: %llu %llu %llu %llu vs %llu %llu %llu %llu", (unsigned long long)np, variant,
            (unsigned long long)c1[0], (unsigned long long)c1[1], (unsigned long long)c1[2], (unsigned long long)c1[3],
            (unsigned long long)c2[0], (unsigned long long)c2[1], (unsigned long long)c2[2], (unsigned long long)c2[3]);
      CHECK(c1[0] > 0 || np < 8 || variant != 0, "derand degenerate input np=%llu", (unsigned long long)np);
    }
  }
}

// ------------------------------------------------------------------------------------------------ k_sync_track
static bool same_result(const SyncResult &x, const SyncResult &y) {
  bool ok = x.st.synchronized == y.st.synchronized && x.st.bitphase == y.st.bitphase && x.st.polarity == y.st.polarity &&
            x.st.phase8 == y.st.phase8 && x.st.next_sync_count == y.st.next_sync_count && x.st.lock_timeleft == y.st.lock_timeleft &&
            x.st.locktime == y.st.locktime && x.st.report_state == y.st.report_state && x.st.fastlock == y.st.fastlock &&
            x.st.resync_period == y.st.resync_period && x.st.resync_phase == y.st.resync_phase && x.consumed == y.consumed &&
            x.produced == y.produced && x.need_next_sync == y.need_next_sync && x.events == y.events;
  for (int i = 0; ok && i < x.events && i < 16; ++i) ok = x.event_val[i] == y.event_val[i] && x.event_pos[i] == y.event_pos[i];
  return ok;
}

static void case_sync_locked(uint64_t seed) {
  std::mt19937_64 rng(seed);
  const uint64_t sizes[] = {0, 1, 31, 32, 33, 64, 1000, 65535, 65536, 65537, 65600, 131072, 200003};
  for (uint64_t np : sizes) {
    for (int variant = 0; variant < 6; ++variant) {
      // 0: all good; 1: one bad packet; 2: sparse bad packets (no 3 in a row, mostly); 3: dense (loses lock early);
      // 4: a burst right after the first tile; 5: bad packets at the very start with a short time left
      std::vector<uint32_t> words((np + 31) / 32 + 2, 0);
      auto setbad = [&](uint64_t p) { if (p < np) words[p >> 5] |= 1u << (p & 31); };
      if (variant == 1 && np) setbad(rng() % np);
      if (variant == 2) for (uint64_t k = 0; k < np / 50 + 1; ++k) setbad(rng() % (np + 1));
      if (variant == 3) for (uint64_t p = 0; p < np; ++p) if (rng() % 3 != 0) setbad(p);
      if (variant == 4) for (uint64_t p = 65536 + rng() % 40; p < 65536 + 80; ++p) setbad(p);
      if (variant == 5) for (uint64_t p = 0; p < 2; ++p) setbad(p);
      SyncState st{};
      st.synchronized = 1; st.bitphase = (int)(rng() % 8); st.polarity = (rng() & 1) ? 0xff : 0; st.phase8 = (int)(rng() % 8);
      st.next_sync_count = (int)(rng() % 3); st.lock_timeleft = 1 + (uint32_t)(rng() % 4); st.locktime = rng() % 100000;
      st.report_state = (variant == 2); st.fastlock = 0; st.resync_period = 32; st.resync_phase = 0;
      if (variant == 5) st.lock_timeleft = 2;
      SyncResult r1, r2;
      memset(&r1, 0xee, sizeof r1); memset(&r2, 0xee, sizeof r2);
      emu::launch(1, 256, [&] { v1::k_sync_track(nullptr, 204 * np, &st, words.data(), np, &r1); });
      emu::launch(1, 256, [&] { v2::k_sync_track(nullptr, 204 * np, &st, words.data(), np, &r2); });
      CHECK(same_result(r1, r2), "sync locked np=%llu variant=%d: consumed %llu vs %llu, sync %d vs %d, timeleft %u vs %u, locktime %llu vs %llu",
            (unsigned long long)np, variant, (unsigned long long)r1.consumed, (unsigned long long)r2.consumed, r1.st.synchronized, r2.st.synchronized,
            r1.st.lock_timeleft, r2.st.lock_timeleft, (unsigned long long)r1.st.locktime, (unsigned long long)r2.st.locktime);
    }
  }
}

static void case_sync_search(uint64_t seed) {
  std::mt19937_64 rng(seed);
  int nlocked = 0;
  for (int variant = 0; variant < 6; ++variant) {
    // a byte stream with MPEG syncs every 204 bytes from a random byte/bit offset (variants 0-3), or noise only (4, 5)
    const uint64_t nbytes = 204 * 8 * (variant == 5 ? 30 : 12) + rng() % 300;
    std::vector<uint8_t> clean(nbytes + 8, 0);
    for (auto &b : clean) b = (uint8_t)rng();
    const int bitphase = (int)(rng() % 8);
    const uint64_t off = rng() % 204;
    const int pol = (variant & 1) ? 0xff : 0;
    std::vector<uint8_t> bytes(nbytes + 8, 0);
    if (variant < 4) {
      for (uint64_t i = off, k = 0; i < nbytes; i += 204, ++k) clean[i] = (k % 8 == 3) ? 0xb8 : 0x47;
      // shift the clean stream right by `bitphase` bits, inverted when pol
      for (uint64_t i = 0; i + 1 < bytes.size(); ++i) {
        const unsigned w = ((unsigned)(clean[i] ^ pol) << 8) | (clean[i + 1] ^ pol);
        bytes[i + 1] = (uint8_t)(w >> (8 - bitphase)) ;
      }
    } else {
      for (auto &b : bytes) b = (uint8_t)rng();
    }
    for (int fast = 0; fast < 2; ++fast) {
      SyncState st{};
      st.synchronized = 0; st.bitphase = (int)(rng() % 8); st.polarity = 0; st.phase8 = 0; st.next_sync_count = (int)(rng() % 3);
      st.lock_timeleft = 0; st.locktime = 0; st.report_state = (variant == 0); st.fastlock = fast; st.resync_period = 4; st.resync_phase = (int)(rng() % 4);
      SyncResult r1, r2;
      memset(&r1, 0xee, sizeof r1); memset(&r2, 0xee, sizeof r2);
      emu::launch(1, 256, [&] { v1::k_sync_track(bytes.data(), nbytes, &st, nullptr, 0, &r1); });
      emu::launch(1, 256, [&] { v2::k_sync_track(bytes.data(), nbytes, &st, nullptr, 0, &r2); });
      CHECK(same_result(r1, r2), "sync search variant=%d fast=%d: consumed %llu vs %llu, sync %d vs %d", variant, fast,
            (unsigned long long)r1.consumed, (unsigned long long)r2.consumed, r1.st.synchronized, r2.st.synchronized);
      nlocked += r1.st.synchronized == 1;
    }
  }
  CHECK(nlocked >= 2, "sync search locked %d times only on clean streams", nlocked);
}

// ------------------------------------------------------------------------------------------------ k_deconv_tiled
static void case_deconv(uint64_t seed) {
  std::mt19937_64 rng(seed);
  const int rates[][2] = {{1, 2}, {4, 6}, {3, 4}, {5, 6}, {7, 8}};   // punctperiod, punctweight (dvb.h:256-276)
  for (int ri = 0; ri < 5; ++ri) {
    for (int variant = 0; variant < 6; ++variant) {
      const int pp = rates[ri][0], pw = rates[ri][1], half = pw / 2;
      const uint64_t nsym = (variant == 0) ? 40 + rng() % 64 : 9000 + rng() % 30000;
      const int mis = variant % 4;
      std::vector<uint32_t> store(nsym + 8);
      for (auto &s : store) s = (uint32_t)rng();
      DeconvArgs a{};
      a.symbols = store.data() + mis;
      a.reg_in = rng();
      const int n_in_choices[] = {0, 2, 30, 64 - pw, 64, 64 - pw};
      a.n_in = n_in_choices[variant];
      a.out_acc = rng(); a.n_out = (variant == 1) ? 0 : (int)(rng() % 8);
      const uint8_t perms[4][4] = {{0, 1, 2, 3}, {2, 0, 3, 1}, {3, 2, 1, 0}, {1, 3, 0, 2}};
      for (int k = 0; k < 4; ++k) a.hyp[k] = perms[rng() % 4][k];
      a.punctperiod = pp; a.punctweight = pw;
      for (int b = 0; b < 8; ++b) { a.deconv[b] = rng(); a.deconv2[b] = rng(); }
      const int64_t k0 = (a.n_in >= 64) ? 0 : (64 - a.n_in) / 2;
      // bytes the host would ask for: every group whose register is complete (K_g <= nsym)
      int64_t ngroups = ((int64_t)nsym - k0) / half + 1;
      if ((int64_t)nsym < k0) ngroups = 0;
      a.nbytes = (uint64_t)((a.n_out + ngroups * pp) / 8);
      if (variant == 5 && a.nbytes > 3000) a.nbytes = 1024 * 2;     // (ends exactly on a tile)
      for (int errmode = 0; errmode < 2; ++errmode) {
        std::vector<uint8_t> o1(a.nbytes + 16, 0x5a), o2(a.nbytes + 16, 0x5a);
        unsigned long long e1 = 0, e2 = 0;
        DeconvArgs a1 = a, a2 = a;
        a1.out = o1.data(); a2.out = o2.data();
        a1.err_out = errmode ? &e1 : nullptr; a2.err_out = errmode ? &e2 : nullptr;
        const unsigned grid = (unsigned)((a.nbytes + v2::kDcBytes - 1) / v2::kDcBytes);
        if (!grid) continue;
        emu::launch(grid, 256, [&] { v1::k_deconv_tiled(a1, nsym); });
        emu::launch(grid, 256, [&] { v2::k_deconv_tiled(a2, nsym); });
        CHECK(o1 == o2, "deconv bytes rate %d/%d variant=%d nbytes=%llu", pp, pw, variant, (unsigned long long)a.nbytes);
        CHECK(e1 == e2, "deconv errors rate %d/%d variant=%d: %llu vs %llu", pp, pw, variant, e1, e2);
        if (!errmode) {
          bool touched = false;
          for (uint64_t i = 0; i < a.nbytes; ++i) touched |= o1[i] != 0x5a;
          CHECK(touched, "deconv wrote nothing");
          CHECK(o2[a.nbytes] == 0x5a, "deconv wrote past the end");
        } else {
          CHECK(e1 > 0, "deconv error count is zero");
        }
      }
    }
  }
}

int main(int argc, char **argv) {
  const std::string which = argc > 1 ? argv[1] : "all";
  const uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 0) : 1;
  if (which == "plan" || which == "all") case_plan(seed);
  if (which == "derand" || which == "all") case_derand(seed);
  if (which == "sync_locked" || which == "all") case_sync_locked(seed);
  if (which == "sync_search" || which == "all") case_sync_search(seed);
  if (which == "deconv" || which == "all") case_deconv(seed);
  if (g_fail) { fprintf(stderr, "%d mismatches (case %s, seed %llu)\n", g_fail, which.c_str(), (unsigned long long)seed); return 1; }
  printf("emu_ctl %s seed %llu: identical\n", which.c_str(), (unsigned long long)seed);
  return 0;
}
