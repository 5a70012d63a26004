// emu_tx.cpp -- TEST INFRASTRUCTURE.  Runs the kernels of the transmit chain (leansdr_b200/csrc/tx.cu: k_tx_tsgen,
// k_tx_rs = randomizer + rs_encoder, k_tx_interleave, k_tx_convol; k_tx_resample = cstln_transmitter + fir_resampler +
// decimator; k_tx_amp2 / k_tx_agc / k_tx_scale = simple_agc, float for float) on the host (cuda_emu.h) against the oracle's
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
  // ---- the float stage: cstln_transmitter (sdr.h:1196-1225) + fir_resampler (dsp.h:290-364) + decimator
  //      (generic.h:238-264), fused in k_tx_resample, then simple_agc (sdr.h:232-279) -- FLOAT FOR FLOAT
  {
    ldvb::Cstln cs = ldvb::make_cstln(LDVB_CSTLN_QPSK, LDVB_FEC12, false);
    orc_cstln *oc = (orc_cstln *)malloc(sizeof(orc_cstln));
    orc_cstln_build(oc, ORC_QPSK, 0);
    std::vector<float2> points(cs.nsymbols);
    for (int k = 0; k < cs.nsymbols; ++k) points[k] = make_float2(0 + cs.sym_re[k], 0 + cs.sym_im[k]);
    const size_t nsym = 2500;
    std::vector<uint8_t> sy(nsym);
    for (auto &v : sy) v = (uint8_t)(rng() & 3);
    std::vector<float2> mapped(nsym);
    orc_tx_map(oc, sy.data(), nsym, reinterpret_cast<float *>(mapped.data()));
    const struct { int interp, decim; const char *power; } cfgs[] = {{6, 5, "37.5"}, {2, 1, "30"}, {12, 5, "43.5"}};
    for (const auto &cf : cfgs) {
      const float amp = dev::tx_amp(cf.power);
      CHECK(amp == orc_tx_amp(cf.power), "tx_amp %s", cf.power);
      const std::vector<float> taps = ldvb::design_tx_rrc(cf.interp, 0.35f, 10.f, amp);
      std::vector<float> otaps(65536);
      const int on = orc_tx_taps(cf.interp, 0.35f, 10.f, amp, otaps.data());
      CHECK(on == (int)taps.size() && memcmp(taps.data(), otaps.data(), 4 * taps.size()) == 0, "RRC taps, interpolation %d", cf.interp);
      const std::vector<float> sc = ldvb::shift_taps_resampler(taps, 0.f);
      std::vector<float2> interp_out(nsym * cf.interp + 16);
      size_t consumed = 0;
      const size_t ni = orc_tx_resample(reinterpret_cast<const float *>(mapped.data()), nsym, otaps.data(), on, cf.interp,
                                        reinterpret_cast<float *>(interp_out.data()), &consumed);
      const size_t nd = ni / cf.decim;                                   // decimator::run: whole groups of `decim`
      CHECK(nd > 1000, "resampler %d/%d: the oracle made %zu samples", cf.interp, cf.decim, nd);
      std::vector<float2> res(nd + 4, make_float2(-1.f, -1.f));
      std::vector<unsigned char> smem(taps.size() * 8 + 64);
      emu::g_dyn_smem = smem.data();
      for (int from_symbols = 0; from_symbols < 2; ++from_symbols) {
        // two launches with a ragged split: m0 carries the absolute output index across pushes
        const uint64_t split = nd / 3 + 7;
        for (int part = 0; part < 2; ++part) {
          dev::TxResampleArgs a{};
          a.sym = sy.data(); a.x = mapped.data(); a.points = points.data(); a.taps = reinterpret_cast<const float2 *>(sc.data());
          a.ncoeffs = (int)taps.size(); a.interp = cf.interp; a.decim = cf.decim; a.latency = (a.ncoeffs + cf.interp) / cf.interp;
          a.n0 = 0; a.m0 = part ? split : 0; a.count = part ? nd - split : split; a.out = res.data() + a.m0;
          if (from_symbols) emu::launch((unsigned)((a.count + 255) / 256), 256, [&] { dev::k_tx_resample<true>(a); });
          else emu::launch((unsigned)((a.count + 255) / 256), 256, [&] { dev::k_tx_resample<false>(a); });
        }
        size_t bad = 0;
        for (size_t m = 0; m < nd; ++m) bad += memcmp(&res[m], &interp_out[m * cf.decim], 8) != 0;
        CHECK(bad == 0 && res[nd].x == -1.f, "resampler %d/%d (%s): %zu of %zu samples differ", cf.interp, cf.decim,
              from_symbols ? "symbols" : "cf32", bad, nd);
      }
      // simple_agc over the decimated stream, in two pushes (the estimate is carried)
      const float out_rms = amp / sqrtf((float)cf.interp / cf.decim), bw = 0.001 * cf.decim / cf.interp;   // leandvbtx.cc:163-165
      std::vector<float2> want(nd), got(nd, make_float2(-1.f, -1.f));
      std::vector<float2> dec(nd);
      for (size_t m = 0; m < nd; ++m) dec[m] = interp_out[m * cf.decim];
      const size_t na = orc_tx_agc(reinterpret_cast<const float *>(dec.data()), nd, out_rms, bw, reinterpret_cast<float *>(want.data()));
      const uint64_t nchunks = nd / 128, c1 = nchunks / 2 + 1;
      CHECK(na == nchunks * 128, "simple_agc: the oracle wrote %zu samples", na);
      std::vector<float> amp2(nchunks), gain(nchunks);
      float est = 0.f;
      for (int part = 0; part < 2; ++part) {
        const uint64_t cb = part ? c1 : 0, cn = part ? nchunks - c1 : c1;
        emu::launch((unsigned)((cn + 127) / 128), 128, [&] { dev::k_tx_amp2(dec.data() + cb * 128, cn, amp2.data() + cb); });
        emu::launch(1, 256, [&] { dev::k_tx_agc(amp2.data() + cb, cn, bw, out_rms, &est, gain.data() + cb); });
        emu::launch((unsigned)((cn * 128 + 255) / 256), 256, [&] { dev::k_tx_scale(dec.data() + cb * 128, gain.data() + cb, cn * 128, got.data() + cb * 128); });
      }
      CHECK(memcmp(got.data(), want.data(), na * 8) == 0, "simple_agc %d/%d", cf.interp, cf.decim);
    }
    free(oc);
  }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_tx seed %llu: equal\n", (unsigned long long)seed);
  return 0;
}
