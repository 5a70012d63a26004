// emu_fec.cpp -- TEST INFRASTRUCTURE.  Runs the packet kernels of leansdr_b200/csrc/k_fec.cu on the host (cuda_emu.h):
// the Reed-Solomon decoder with and without the fused de-interleaver gather (k_rs), the byte re-alignment (k_realign),
// the sync flags (k_sync_flags) and the de-randomiser's output grid behind its scan (k_derand_*).  The device text is
// cut out of k_fec.cu by the test (everything inside its anonymous namespace; the launchers stay behind) and included
// here as FEC_DEV_INC.  The checker is the oracle's C restatement (liboracle.so): RS packets, flags and corrected-bit
// counts, de-randomised TS.  Built with -fsanitize=thread the same run is the race check of these kernels.
// Usage: emu_fec <case: rs|rs_deint|realign|derand> <seed>; exit code 0 = equal.
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

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: emu_fec <rs|rs_deint|realign|derand> <seed>\n"); return 2; }
  const std::string c = argv[1];
  const uint64_t seed = strtoull(argv[2], nullptr, 10);
  if (c == "rs") case_rs(seed, false);
  else if (c == "rs_deint") case_rs(seed, true);
  else if (c == "realign") case_realign(seed);
  else if (c == "derand") case_derand(seed);
  else { fprintf(stderr, "unknown case\n"); return 2; }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_fec %s seed %llu: equal\n", c.c_str(), (unsigned long long)seed);
  return 0;
}
