// emu_tx.cpp -- TEST INFRASTRUCTURE.  Runs the integer kernels of the transmit chain (leansdr_b200/csrc/tx.cu: k_tx_tsgen,
// k_tx_rs = randomizer + rs_encoder, k_tx_interleave, k_tx_convol) on the host (cuda_emu.h) against the oracle's
// restatement of leandvbtx (oracle/dvbs_tx_oracle.c, pinned to the reference transmitter's digests).  The device text
// is cut out of tx.cu by the test (its first anonymous namespace: kernels and the two host table builders) and
// included as TX_DEV_INC.  Usage: emu_tx <seed>; exit code 0 = equal.
#include "cuda_emu.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../leansdr_b200/csrc/kernels.h"
#include "../../leansdr_b200/csrc/tables.h"
#include "../../include/leandvb_b200.h"
extern "C" {
#include "../../oracle/dvbs_oracle.h"
}

inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
namespace ldvb {   // (what common.cuh gives the device code; built with -ffp-contract=off)
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
inline float2 cmul(float2 a, float2 b) { return make_float2(fsub(fmul(a.x, b.x), fmul(a.y, b.y)), fadd(fmul(a.x, b.y), fmul(a.y, b.x))); }
}  // namespace ldvb

namespace dev {
#include TX_DEV_INC
}

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

int main(int argc, char **argv) {
  const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  std::mt19937_64 rng(seed);
  const uint64_t npk = 30, first = 5 + rng() % 70000;
  // ---- leantsgen (apps/leantsgen.cc:37-47)
  std::vector<uint8_t> ts(188 * npk, 0xee);
  emu::launch((unsigned)((npk * 47 + 255) / 256), 256, [&] { dev::k_tx_tsgen(first, npk, ts.data()); });
  for (uint64_t p = 0; p < npk; ++p) {
    const uint32_t t = (uint32_t)(first + p);
    for (int k = 0; k < 47; ++k) {
      const uint8_t want[4] = {(uint8_t)(k ? 4 * k : 0x47), (uint8_t)(t >> 16), (uint8_t)(t >> 8), (uint8_t)t};
      CHECK(memcmp(&ts[188 * p + 4 * k], want, 4) == 0, "tsgen packet %llu word %d", (unsigned long long)p, k);
    }
  }
  for (auto &b : ts) if (rng() % 3 == 0) b = (uint8_t)rng();            // (arbitrary payloads from here on)
  for (uint64_t p = 0; p < npk; ++p) ts[188 * p] = 0x47;
  // ---- randomizer + rs_encoder (dvb.h:1073-1095, 957-980); packet 0 of the run sits at PRBS position 0
  uint8_t ex[512], lg[256], G[17];
  ldvb::make_rs_tables(ex, lg);
  dev::tx_rs_generator(ex, lg, G);
  const std::vector<uint8_t> pattern = ldvb::make_derand_pattern();
  std::vector<uint8_t> rs(204 * npk, 0xee);
  dev::TxRsArgs ra{};
  ra.ts = ts.data(); ra.rs = rs.data(); ra.first_packet = 0; ra.n = (uint32_t)npk; ra.pattern = pattern.data();
  ra.gf_exp = ex; ra.gf_log = lg; ra.g = G;
  emu::launch((unsigned)((npk + 3) / 4), 128, [&] { dev::k_tx_rs(ra); });
  std::vector<uint8_t> rnd(188 * npk), want_rs(204 * npk);
  orc_tx_randomize(ts.data(), npk, rnd.data());
  orc_tx_rs_encode(rnd.data(), npk, want_rs.data());
  CHECK(rs == want_rs, "randomizer + rs_encoder");
  // ---- interleaver (dvb.h:896-918)
  const uint64_t rows = npk - 11;
  std::vector<uint8_t> mb(2 + rows * 204, 0), want_mb(rows * 204 + 16);
  emu::launch((unsigned)((rows * 204 + 255) / 256), 256, [&] { dev::k_tx_interleave(rs.data(), rows, mb.data() + 2); });
  const size_t nmb = orc_tx_interleave(want_rs.data(), npk, want_mb.data());
  CHECK(nmb == rows * 204 && memcmp(mb.data() + 2, want_mb.data(), nmb) == 0, "interleaver");
  // ---- dvb_convol for every code rate and symbol width it takes (dvb.h:519-604, convolutional.h:225-270);
  //      the stream starts with an all-zero history: two zero bytes in front
  for (int fec = 0; fec <= 5; ++fec) {
    for (int bps = 1; bps <= 6; ++bps) {
      int bits_in = 0, bits_out = 0; uint16_t polys[8] = {0};
      if (!dev::tx_fec_spec(fec, &bits_in, &bits_out, polys) || bits_out % bps) continue;
      const uint64_t nbytes = (rows * 204 / bits_in) * bits_in;                      // dvb.h:591-593
      const uint64_t ngroups = nbytes * 8 / bits_in;
      std::vector<uint8_t> sym(ngroups * bits_out / bps + 8, 0xee), want(sym.size() + 64);
      dev::TxConvArgs ca{};
      ca.bytes = mb.data(); ca.ngroups = ngroups; ca.bits_in = bits_in; ca.bits_out = bits_out; ca.bps = bps;
      memcpy(ca.polys, polys, sizeof polys);
      ca.sym = sym.data();
      emu::launch((unsigned)((ngroups + 255) / 256), 256, [&] { dev::k_tx_convol(ca); });
      size_t consumed = 0;
      const size_t nw = orc_tx_convol(fec, bps, mb.data() + 2, rows * 204, want.data(), &consumed);
      CHECK(consumed == nbytes && nw == ngroups * bits_out / bps, "convol fec %d bps %d: oracle consumed %zu wrote %zu", fec, bps, consumed, nw);
      CHECK(memcmp(sym.data(), want.data(), nw) == 0, "convol fec %d bps %d: symbols", fec, bps);
      CHECK(sym[nw] == 0xee, "convol fec %d bps %d: wrote past the end", fec, bps);
    }
  }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_tx seed %llu: equal\n", (unsigned long long)seed);
  return 0;
}
