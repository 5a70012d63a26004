// emu_rx.cpp -- TEST INFRASTRUCTURE.  Runs the receiver kernels of leansdr_b200/csrc/k_rx.cu -- k_rx_serial (EXACT
// mode, settling pass), k_rx (one lane per time span: warm-up, owned chunks, verification overlap, seam logs; repair
// mode re-runs a span from its predecessor's end state), k_rx_stitch, k_rx_plan_local/_apply and k_rx_compact -- on the
// host (cuda_emu.h) against the oracle's cstln_receiver (orc_rx_run, sdr.h:772-915), FIELD FOR FIELD: this file is built
// with -ffp-contract=off and the shim's fmul / fadd / cmul are single IEEE operations like the device's _rn
// intrinsics.  The device text is the anonymous namespace of k_rx.cu (RX_DEV_INC: its three inline-PTX statements
// replaced by their C meaning, the dynamic shared memory pointed at the shim's buffer); asynchronous row copies are
// done at issue time.  Launch geometry: two warps per CTA instead of twelve (the row windows of the warps are laid out
// the same way).  Built with -fsanitize=thread the same run is the kernels' race check (rows copied by other lanes,
// __syncwarp on both sides of a stage).  Usage: emu_rx <seed> [quick]; exit code 0 = equal.
#include "cuda_emu.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

extern "C" {
#include "../../oracle/dvbs_oracle.h"
}

inline double __dmul_rn(double a, double b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xffu) << (8 * i);
  return r;
}
template <class T> inline void __stcg(T *p, T v) { *p = v; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
template <class T> inline T __ldcs(const T *p) { return *p; }
#include "../../leansdr_b200/csrc/kernels.h"
#include "../../leansdr_b200/csrc/tables.h"
namespace ldvb {
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
inline float2 cmul(float2 a, float2 b) { return make_float2(fsub(fmul(a.x, b.x), fmul(a.y, b.y)), fadd(fmul(a.x, b.y), fmul(a.y, b.x))); }
inline int f2i_trunc(float a) { return (int)a; }
inline uint32_t smem_u32(const void *p) { return (uint32_t)(reinterpret_cast<const unsigned char *>(p) - emu::g_dyn_smem); }
inline void cp_async16(void *dst, const void *src) { memcpy(dst, src, 16); }
inline void cp_async16_ca(void *dst, const void *src) { memcpy(dst, src, 16); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
namespace dev {
#include RX_DEV_INC
}
}  // namespace ldvb
using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// A QPSK-like waveform at `omega` samples per symbol: linear ramps between the points, a slow carrier rotation, noise,
// short bursts.
static std::vector<float2> waveform(std::mt19937_64 &rng, size_t n, float omega, float amp, float cfo, float noise, bool bursts = true, bool shaped = false) {
  std::vector<float2> x(n);
  std::normal_distribution<float> g(0.f, noise);
  const size_t nsym = (size_t)(n / omega) + 4;
  std::vector<float2> s(nsym);
  for (auto &v : s) { const unsigned b = (unsigned)(rng() & 3); v = make_float2((b & 2) ? -amp : amp, (b & 1) ? -amp : amp); }
  for (size_t i = 0; i < n; ++i) {
    const double t = (i + 0.37) / omega;
    const size_t k = (size_t)t;
    float f = (float)(t - k);
    if (shaped) f = f < 0.6f ? 0.f : (f - 0.6f) / 0.4f;   // plateaus with short transitions (little inter-symbol interference)
    const float re = s[k].x * (1 - f) + s[k + 1].x * f, im = s[k].y * (1 - f) + s[k + 1].y * f;
    const double a = 2 * M_PI * cfo * (double)i + 0.4;
    const float c = (float)cos(a), sn = (float)sin(a);
    x[i] = make_float2(re * c - im * sn + g(rng), re * sn + im * c + g(rng));
    if (bursts && i % 4099 < 3) { x[i].x *= 7.f; x[i].y *= 7.f; }   // bursts: the slicer's halving loop (sdr.h:476-481) and the mu clamp
  }
  return x;
}

struct Tables {
  Cstln cst;
  std::vector<float> trig;
  std::vector<int16_t> pe;
  std::vector<uint8_t> rot_perm;
  orc_cstln *oc;
  std::vector<float> otrig;
};

static Tables make_tables(int kind) {
  Tables t;
  t.cst = make_cstln(kind, 3 /* 3/4: the APSK ring ratios orc_cstln_build uses */, false);
  t.trig = make_trig16();
  t.oc = (orc_cstln *)malloc(sizeof(orc_cstln));
  orc_cstln_build(t.oc, kind, 0);
  t.otrig.resize(65536 * 2);
  orc_trig16_build(t.otrig.data());
  // the folded phase-error column of slicer 1, as ldvb_create lays it out (pipeline.cu)
  t.pe.assign((size_t)256 * kPeFoldPitch, 0);
  for (int ib = 0; ib < 256; ++ib)
    for (int q = 0; q <= 128; ++q) {
      const int qb = (q == 128) ? 0x80 : q;
      const int v = t.cst.cells[(size_t)ib * 256 + qb].phase_error;
      t.pe[(size_t)ib * kPeFoldPitch + q] = (int16_t)((q == 128) ? -v : v);
    }
  for (int k = 0; k < t.cst.nrotations; ++k)
    for (int s = 0; s < t.cst.nsymbols; ++s) t.rot_perm.push_back(t.cst.rot[k][s]);
  return t;
}

struct OracleOut {
  std::vector<uint32_t> sym;
  std::vector<float2> sampled;
  std::vector<float> meas;
  orc_rx r;
};

static void compare_state(const RxState &s, const orc_rx &r, int sampler, const char *what) {
  CHECK(fbits(s.mu) == fbits(r.mu) && fbits(s.phase) == fbits(r.phase) && fbits(s.freqw) == fbits(r.freqw), "%s: mu/phase/freqw %g %g %g vs %g %g %g",
        what, s.mu, s.phase, s.freqw, r.mu, r.phase, r.freqw);
  CHECK(fbits(s.est_insp) == fbits(r.est_insp) && fbits(s.agc_gain) == fbits(r.agc_gain), "%s: AGC %g %g vs %g %g", what, s.est_insp, s.agc_gain, r.est_insp, r.agc_gain);
  CHECK(fbits(s.est_sp) == fbits(r.est_sp) && fbits(s.est_ep) == fbits(r.est_ep), "%s: estimators", what);
  for (int k = 0; k < 3; ++k)
    CHECK(fbits(s.hist[4 * k]) == fbits(r.hist[k].p_re) && fbits(s.hist[4 * k + 1]) == fbits(r.hist[k].p_im) &&
          fbits(s.hist[4 * k + 2]) == fbits(r.hist[k].c_re) && fbits(s.hist[4 * k + 3]) == fbits(r.hist[k].c_im), "%s: hist[%d]", what, k);
  CHECK(fbits(s.freq_tap) == fbits(r.freq_tap) && s.meas_count == (uint32_t)r.meas_count, "%s: freq_tap / meas_count", what);
  if (sampler == 1) CHECK(fbits(s.samp_freqw) == fbits(r.samp_freqw), "%s: samp_freqw", what);
  if (sampler == 2) CHECK(s.rrc_update_phase == r.rrc_update_phase, "%s: rrc_update_phase %d vs %d", what, s.rrc_update_phase, r.rrc_update_phase);
}

template <int SLICER>
static void run_case(std::mt19937_64 &rng, const Tables &t, int sampler, float omega, uint32_t nspans, uint32_t S, uint32_t W,
                     const char *name) {
  const uint64_t nchunks = (uint64_t)nspans * S - 1;          // ragged last span
  std::vector<float> rrc;
  int rrc_sub = 0;
  if (sampler == 2) { int steps = 0; rrc = design_rrc(2.4e6f * omega / 1.2f, 2.0e6f, 0.35f, 10.f, 0, &steps); rrc_sub = steps; }
  const size_t look = sampler == 2 ? rrc.size() + 8 : 8;
  const std::vector<float2> x = waveform(rng, (size_t)nchunks * kRxChunk + look, omega, 38.f, 0.0013f, 2.5f);

  // ---- oracle: one serial pass
  OracleOut o;
  orc_rx_init(&o.r, t.oc, t.otrig.data(), sampler);
  orc_rx_set_omega(&o.r, omega);
  o.r.meas_decimation = 1000;
  if (sampler == 2) orc_rx_set_rrc(&o.r, (int)rrc.size(), rrc.data(), rrc_sub);
  RxState st0;
  memset(&st0, 0, sizeof st0);
  st0.est_insp = o.r.est_insp; st0.agc_gain = o.r.agc_gain; st0.freqw = o.r.freqw; st0.freq_tap = o.r.freq_tap;
  {
    o.sym.resize((size_t)nchunks * kRxChunk);
    o.sampled.resize(nchunks);
    o.meas.resize(3 * (nchunks + 1));
    size_t ns = 0, np = 0, nm = 0;
    const size_t n_in = (size_t)nchunks * kRxChunk + (size_t)orc_rx_readahead(&o.r);
    const size_t done = orc_rx_run(&o.r, reinterpret_cast<const float *>(x.data()), n_in, reinterpret_cast<uint8_t *>(o.sym.data()), &ns,
                                   reinterpret_cast<float *>(o.sampled.data()), &np, o.meas.data(), &nm);
    CHECK(done == (size_t)nchunks * kRxChunk, "%s: the oracle consumed %zu samples", name, done);
    o.sym.resize(ns); o.sampled.resize(np); o.meas.resize(3 * nm);
  }

  // ---- product parameters (rx_setup, pipeline.cu)
  RxArgs a;
  memset(&a, 0, sizeof a);
  RxParams &p = a.p;
  p.cstln = reinterpret_cast<const CstlnCellDev *>(t.cst.cells.data());
  p.trig = reinterpret_cast<const float2 *>(t.trig.data());
  for (int s = 0; s < t.cst.nsymbols; ++s) { p.sym_re[s] = t.cst.sym_re[s]; p.sym_im[s] = t.cst.sym_im[s]; }
  p.nsymbols = t.cst.nsymbols; p.sampler = sampler; p.omega = omega;
  p.min_freqw = o.r.min_freqw; p.max_freqw = o.r.max_freqw;
  p.freq_alpha = 0.04; p.freq_beta = 0.0012 / omega * 1.0f; p.gain_mu = 0.02 / (75.0f * 75.0f) * 2; p.kest = 0.01f;
  p.allow_drift = 0; p.meas_decimation = 1000;
  p.rrc_coeffs = rrc.data(); p.rrc_n = (int)rrc.size(); p.rrc_sub = rrc_sub;
  p.pe16 = t.pe.data(); p.slicer = SLICER;
  a.x = x.data(); a.nchunks = nchunks; a.avail_chunks = nchunks; a.chunk0 = 0; a.first_exact = 1;
  a.state_in = &st0; a.warm_in = &st0; a.state_chunk = 0;

  auto words_equal = [&](const uint32_t *got, size_t n, size_t at, const char *what) {
    size_t bad = 0;
    for (size_t i = 0; i < n && at + i < o.sym.size(); ++i) bad += (got[i] & 0xffffffu) != (o.sym[at + i] & 0xffffffu);
    CHECK(bad == 0 && at + n <= o.sym.size(), "%s %s: %zu of %zu softsymbols differ (oracle has %zu, span at %zu)", name, what, bad, n, o.sym.size(), at);
  };

  const size_t pe_bytes = SLICER ? (size_t)256 * kPeFoldPitch * 2 : 0;
  // ---- 1. the serial lane: one span over the batch (EXACT mode)
  {
    const uint32_t cap = (uint32_t)(((size_t)nchunks * kRxChunk + 3) & ~(size_t)3);
    std::vector<uint32_t> out(cap + 4, 0xdeadbeefu);
    RxSpanInfo info; RxState end, begin;
    std::vector<float2> sampled(nchunks); std::vector<uint32_t> flag(nchunks, 7u);
    std::vector<float> meas(4 * (nchunks + 1)); uint32_t nmeas = 0;
    RxArgs b = a;
    b.span_chunks = (uint32_t)nchunks; b.warm_chunks = 0; b.nspans = 1; b.span_cap = cap;
    b.sym_out = out.data(); b.info = &info; b.state_end = &end; b.state_begin = &begin;
    b.sampled = sampled.data(); b.sampled_flag = flag.data(); b.meas = meas.data(); b.meas_count = &nmeas; b.max_meas = (uint32_t)nchunks + 1;
    std::vector<unsigned char> smem(pe_bytes + 2 * (kRxChunk + 8) * 8 + 256);
    emu::g_dyn_smem = smem.data();
    emu::launch(1, SLICER ? 128 : 32, [&] { dev::k_rx_serial<SLICER>(b); });
    CHECK(info.n_out == o.sym.size(), "%s serial: %u symbols, oracle %zu", name, info.n_out, o.sym.size());
    words_equal(out.data(), info.n_out, 0, "serial");
    compare_state(end, o.r, sampler, "serial end state");
    size_t k = 0;
    for (uint64_t c = 0; c < nchunks; ++c)
      if (flag[c]) { CHECK(k < o.sampled.size() && fbits(sampled[c].x) == fbits(o.sampled[k].x) && fbits(sampled[c].y) == fbits(o.sampled[k].y), "%s: sampled tap of chunk %llu", name, (unsigned long long)c); ++k; }
    CHECK(k == o.sampled.size(), "%s: %zu sampled points, oracle %zu", name, k, o.sampled.size());
    CHECK(nmeas == o.meas.size() / 3, "%s: %u measurement rows, oracle %zu", name, nmeas, o.meas.size() / 3);
    for (uint32_t m = 0; m < nmeas && m < o.meas.size() / 3; ++m) {
      const float mer = meas[4 * m + 3] >= 0 ? 10 * logf(meas[4 * m + 3]) / logf(10) : 0;     // the host side of the row (sdr.h:910-911)
      CHECK(fbits(meas[4 * m + 1]) == fbits(o.meas[3 * m]) && fbits(meas[4 * m + 2]) == fbits(o.meas[3 * m + 1]) && fbits(mer) == fbits(o.meas[3 * m + 2]),
            "%s: measurement row %u", name, m);
    }
  }

  // ---- 2. the span kernel: every span at once (speculative), then spans 1.. re-run from their predecessors' end states
  const uint32_t cap = (S + kRxVerifyChunks) * kRxChunk;          // a multiple of 4: symbols leave four at a time
  std::vector<uint32_t> out((size_t)nspans * cap, 0xdeadbeefu);
  std::vector<RxSpanInfo> info(nspans);
  std::vector<RxState> end(nspans), begin(nspans);
  std::vector<RxSeamSym> hlog((size_t)nspans * kRxSeamLog), tlog((size_t)nspans * kRxSeamLog);
  CHECK((reinterpret_cast<uintptr_t>(out.data()) & 15u) == 0 && cap % 4 == 0, "%s: the 16-byte emission path is not the one under test", name);
  RxArgs b = a;
  b.span_chunks = S; b.warm_chunks = W; b.nspans = nspans; b.span_cap = cap;
  b.sym_out = out.data(); b.info = info.data(); b.state_end = end.data(); b.state_begin = begin.data();
  b.head_log = hlog.data(); b.tail_log = tlog.data();
  const int warps = 2;
  const size_t row_bytes = sampler == 2 ? (8 + 6) * 8 : (8 + 2) * 8;
  std::vector<unsigned char> smem(pe_bytes + (size_t)warps * dev::kStages * 32 * row_bytes + 256);
  emu::g_dyn_smem = smem.data();
  emu::launch((nspans + warps * 32 - 1) / (warps * 32), warps * 32, [&] { dev::k_rx<8, SLICER>(b, nullptr, 0); });
  // span 0 started from the true state: its symbols are the oracle's, and so is its end state's continuation
  words_equal(out.data(), info[0].n_out, 0, "span 0");
  for (uint32_t j = 1; j < nspans; ++j) {
    const uint32_t list = j;
    emu::launch(1, warps * 32, [&] { dev::k_rx<8, SLICER>(b, &list, 1); });
  }
  size_t at = 0;
  for (uint32_t j = 0; j < nspans; ++j) {
    CHECK(info[j].n_out + info[j].n_tail <= cap, "%s: span %u overflows", name, j);
    words_equal(out.data() + (size_t)j * cap, info[j].n_out, at, "repaired span");
    if (j + 1 < nspans) words_equal(out.data() + (size_t)j * cap + info[j].n_out, info[j].n_tail, at + info[j].n_out, "verification overlap");
    at += info[j].n_out;
  }
  CHECK(at == o.sym.size(), "%s: spans hold %zu symbols, oracle %zu", name, at, o.sym.size());
  compare_state(end[nspans - 1], o.r, sampler, "last span end state");

  // ---- 3. seams of the exact spans verify under the strict rule, the plan and the compaction give the oracle's stream
  std::vector<RxSeam> seams(nspans);
  RxStitchArgs sa;
  memset(&sa, 0, sizeof sa);
  sa.info = info.data(); sa.head_log = hlog.data(); sa.tail_log = tlog.data(); sa.nspans = nspans;
  sa.nrot = t.cst.nrotations; sa.nsymbols = t.cst.nsymbols; sa.rot_perm = t.rot_perm.data(); sa.omega = omega;
  sa.seams = seams.data(); sa.strict = 1; sa.state_begin = begin.data(); sa.state_end = end.data();
  sa.tol_phase = 1.f; sa.tol_freqw = 1e-3f;
  emu::launch((nspans - 1 + 3) / 4, 128, [&] { dev::k_rx_stitch(sa, nullptr, 0); });
  for (uint32_t j = 0; j + 1 < nspans; ++j)
    CHECK(seams[j].ok == 1 && seams[j].rot == 0 && seams[j].mismatches == 0 && seams[j].compared >= 8 && seams[j].extend_prev == 0 &&
          seams[j].skip_next == 0 && seams[j].dphase == 0.f && seams[j].dfreqw == 0.f,
          "%s: seam %u ok %d rot %d mism %d of %d extend %d skip %d dphase %g dfreqw %g", name, j, seams[j].ok, seams[j].rot, seams[j].mismatches,
          seams[j].compared, seams[j].extend_prev, seams[j].skip_next, seams[j].dphase, seams[j].dfreqw);
  {
    // the seam between two ranks of the time-sharded mode (k_rx_stitch_pair: the earlier span's tail log and end state are
    // imported): the same judgement as k_rx_stitch's on seam 0
    RxStitchArgs sp = sa;
    sp.info = info.data() + 1; sp.head_log = hlog.data() + kRxSeamLog; sp.state_begin = begin.data() + 1;
    RxSeam pair;
    memset(&pair, 0xff, sizeof pair);
    emu::launch(1, 32, [&] { dev::k_rx_stitch_pair(sp, tlog.data(), info[0].n_tail, &end[0], &pair); });
    CHECK(memcmp(&pair, &seams[0], sizeof pair) == 0, "%s: k_rx_stitch_pair differs from k_rx_stitch on seam 0", name);
    // mean power of the first samples (what decides the cold-start settling pass)
    float pw = -1.f;
    const uint32_t npw = (uint32_t)std::min<size_t>(3000, x.size());
    emu::launch(1, 256, [&] { dev::k_rx_power(x.data(), npw, &pw); });
    double ref = 0;
    for (uint32_t i = 0; i < npw; ++i) ref += (double)x[i].x * x[i].x + (double)x[i].y * x[i].y;
    ref /= npw;
    CHECK(fabs(pw - ref) <= 1e-4 * ref, "%s: k_rx_power %g vs %g", name, pw, ref);
  }
  const unsigned tile = 64;                                      // (1024 spans per CTA in the library)
  const uint32_t nblk = (nspans + tile - 1) / tile;
  std::vector<uint64_t> span_offset((size_t)nspans + 1 + 2 * nblk + 4, 0);
  std::vector<uint32_t> span_skip(nspans, 0);
  std::vector<uint8_t> span_rot(nspans, 0);
  unsigned long long result[9] = {0};
  unsigned long long *totals = reinterpret_cast<unsigned long long *>(span_offset.data() + nspans + 1);
  emu::launch(nblk, tile, [&] { dev::k_rx_plan_local(info.data(), seams.data(), nspans, cap, sa.nrot, 0, 0, span_offset.data(), span_skip.data(), span_rot.data(), totals, result); });
  emu::launch(nblk, tile, [&] { dev::k_rx_plan_apply(nspans, sa.nrot, span_offset.data(), span_rot.data(), totals, result); });
  CHECK(result[0] == 0 && result[1] == o.sym.size() && result[3] == 0, "%s: plan: %llu failed seams, %llu symbols kept (oracle %zu), %llu overflows", name,
        result[0], result[1], o.sym.size(), result[3]);
  std::vector<uint32_t> flat(o.sym.size() + 8, 0xdeadbeefu);
  RxCompactArgs ca;
  memset(&ca, 0, sizeof ca);
  ca.sym_in = out.data(); ca.span_cap = cap; ca.nspans = nspans; ca.span_offset = span_offset.data(); ca.span_skip = span_skip.data();
  ca.span_rot = span_rot.data(); ca.rot_perm = t.rot_perm.data(); ca.nsymbols = t.cst.nsymbols; ca.sym_out = flat.data();
  emu::launch(nspans, 64, [&] { dev::k_rx_compact(ca, result[1]); });
  if (result[1] == o.sym.size()) words_equal(flat.data(), o.sym.size(), 0, "compacted stream");
  CHECK(flat[o.sym.size()] == 0xdeadbeefu, "%s: the compaction wrote past its end", name);
  fprintf(stderr, "%s: %zu symbols over %llu chunks, %u spans, equal so far: %s\n", name, o.sym.size(), (unsigned long long)nchunks, nspans, g_fail ? "NO" : "yes");
}

// --hs: fast_qpsk_receiver<u8> (sdr.h:946-1189), sampler kRxSamplerHs of the same kernels.  Integer loop state in the float
// fields of RxState; the input is the u8 stream as the front end converts it (value - 128).
static void run_case_hs(std::mt19937_64 &rng, float omega, uint32_t nspans, uint32_t S, uint32_t W, const char *name) {
  const uint64_t nchunks = (uint64_t)nspans * S - 1;
  const size_t n = (size_t)nchunks * kRxChunk + 8;
  const std::vector<float2> w = waveform(rng, n, omega, 60.f, 0.0009f, 3.f);
  std::vector<uint8_t> u8(2 * n);
  std::vector<float2> x(n);
  for (size_t i = 0; i < n; ++i) {
    auto q = [](float v) { const int k = (int)lrintf(v) + 128; return (uint8_t)(k < 0 ? 0 : k > 255 ? 255 : k); };
    u8[2 * i] = q(w[i].x); u8[2 * i + 1] = q(w[i].y);
    x[i] = make_float2((float)((int)u8[2 * i] - 128), (float)((int)u8[2 * i + 1] - 128));
  }
  orc_hsrx *r = (orc_hsrx *)malloc(sizeof(orc_hsrx));
  orc_hsrx_init(r);
  orc_hsrx_set_omega(r, omega);
  orc_hsrx_config(r, 0, 1000);
  const long lo = r->min_freqw, hi = r->max_freqw;
  std::vector<uint8_t> osym((size_t)nchunks * kRxChunk);
  std::vector<float> ofreq(nchunks + 1);
  size_t ns = 0, nf = 0;
  const size_t done = orc_hsrx_run(r, u8.data(), (size_t)nchunks * kRxChunk + 1, osym.data(), &ns, ofreq.data(), &nf);
  CHECK(done == (size_t)nchunks * kRxChunk, "%s: the oracle consumed %zu samples", name, done);

  const HsTables ht = make_hs_tables();
  RxState st0;
  memset(&st0, 0, sizeof st0);
  RxArgs a;
  memset(&a, 0, sizeof a);
  RxParams &p = a.p;
  p.nsymbols = 4; p.sampler = kRxSamplerHs; p.omega = omega; p.min_freqw = (float)lo; p.max_freqw = (float)hi;
  p.gain_mu = 0.02 / (75.0f * 75.0f) * 2; p.kest = 0.01f; p.allow_drift = 0; p.meas_decimation = 1000;
  p.hs_polar = ht.polar.data(); p.hs_rect = ht.rect.data(); p.hs_sincos = ht.sincos.data();
  p.hs_freq_beta = (long long)(signed long)(0.0012 * 256 * 65536 / omega * 1.0f);
  a.x = x.data(); a.nchunks = nchunks; a.avail_chunks = nchunks; a.first_exact = 1; a.state_in = &st0; a.warm_in = &st0;

  auto syms_equal = [&](const uint32_t *got, size_t cnt, size_t at, const char *what) {
    size_t bad = 0;
    for (size_t i = 0; i < cnt && at + i < ns; ++i) bad += ((got[i] >> 16) & 0xffu) != osym[at + i];
    CHECK(bad == 0 && at + cnt <= ns, "%s %s: %zu of %zu hard symbols differ (oracle has %zu, span at %zu)", name, what, bad, cnt, ns, at);
  };
  auto state_equal = [&](const RxState &s, const char *what) {
    CHECK(fbits(s.mu) == fbits(r->mu) && s.phase == (float)r->phase && s.freqw == (float)r->freqw && s.meas_count == (uint32_t)r->meas_count,
          "%s %s: mu %g phase %g freqw %g count %u vs %g %u %ld %lu", name, what, s.mu, s.phase, s.freqw, s.meas_count, r->mu, r->phase, r->freqw, r->meas_count);
    for (int k = 0; k < 3; ++k)
      CHECK(s.hist[4 * k] == (float)r->hist[k].p_re && s.hist[4 * k + 1] == (float)r->hist[k].p_im && s.hist[4 * k + 2] == (float)r->hist[k].c_re &&
            s.hist[4 * k + 3] == (float)r->hist[k].c_im, "%s %s: hist[%d]", name, what, k);
  };
  {
    const uint32_t cap = (uint32_t)(((size_t)nchunks * kRxChunk + 3) & ~(size_t)3);
    std::vector<uint32_t> out(cap + 4, 0xdeadbeefu);
    RxSpanInfo info; RxState end;
    std::vector<float> meas(4 * (nchunks + 1)); uint32_t nmeas = 0;
    RxArgs b = a;
    b.span_chunks = (uint32_t)nchunks; b.nspans = 1; b.span_cap = cap; b.sym_out = out.data(); b.info = &info; b.state_end = &end;
    b.meas = meas.data(); b.meas_count = &nmeas; b.max_meas = (uint32_t)nchunks + 1;
    std::vector<unsigned char> smem(2 * (kRxChunk + 8) * 8 + 256);
    emu::g_dyn_smem = smem.data();
    emu::launch(1, 32, [&] { dev::k_rx_serial<0>(b); });
    CHECK(info.n_out == ns, "%s serial: %u symbols, oracle %zu", name, info.n_out, ns);
    syms_equal(out.data(), info.n_out, 0, "serial");
    state_equal(end, "serial end state");
    CHECK(nmeas == nf, "%s: %u frequency rows, oracle %zu", name, nmeas, nf);
    for (uint32_t m = 0; m < nmeas && m < nf; ++m) CHECK(fbits(meas[4 * m + 1]) == fbits(ofreq[m]), "%s: frequency row %u", name, m);
  }
  const uint32_t cap = (S + kRxVerifyChunks) * kRxChunk;
  std::vector<uint32_t> out((size_t)nspans * cap, 0xdeadbeefu);
  std::vector<RxSpanInfo> info(nspans);
  std::vector<RxState> end(nspans);
  std::vector<RxSeamSym> hlog((size_t)nspans * kRxSeamLog), tlog((size_t)nspans * kRxSeamLog);
  RxArgs b = a;
  b.span_chunks = S; b.warm_chunks = W; b.nspans = nspans; b.span_cap = cap; b.sym_out = out.data(); b.info = info.data();
  b.state_end = end.data(); b.head_log = hlog.data(); b.tail_log = tlog.data();
  const int warps = 2;
  std::vector<unsigned char> smem((size_t)warps * dev::kStages * 32 * (8 + 2) * 8 + 256);
  emu::g_dyn_smem = smem.data();
  emu::launch((nspans + warps * 32 - 1) / (warps * 32), warps * 32, [&] { dev::k_rx<8, 0>(b, nullptr, 0); });
  for (uint32_t j = 1; j < nspans; ++j) {
    const uint32_t list = j;
    emu::launch(1, warps * 32, [&] { dev::k_rx<8, 0>(b, &list, 1); });
  }
  size_t at = 0;
  for (uint32_t j = 0; j < nspans; ++j) {
    syms_equal(out.data() + (size_t)j * cap, info[j].n_out, at, "repaired span");
    at += info[j].n_out;
  }
  CHECK(at == ns, "%s: spans hold %zu symbols, oracle %zu", name, at, ns);
  state_equal(end[nspans - 1], "last span end state");
  free(r);
  fprintf(stderr, "%s: %zu symbols over %llu chunks, %u spans, equal so far: %s\n", name, ns, (unsigned long long)nchunks, nspans, g_fail ? "NO" : "yes");
}

// FAST mode as rx_fast_launch / run_receiver (pipeline.cu) schedule it: a serial settling pass, then every span at once from
// the carried loop state (frequency, AGC) after 4 warm-up chunks, strict seams (every hard decision of the overlap and the
// loop states agree), failed seams repaired by an exact re-run of the later span.  The claim of that mode, checked here on the
// kernels themselves: the stitched stream carries the ORACLE's hard decisions (the soft costs may differ by the AGC state
// a span started from), and with every span repaired it is the oracle's stream word for word (run_case above).
static void run_fast_case(std::mt19937_64 &rng, const Tables &t, float noise, bool shaped, size_t max_flips, uint32_t settle, uint32_t nspans, uint32_t S, const char *name) {
  const float omega = 1.2f;
  const uint32_t W = 4;
  const uint64_t nchunks = settle + (uint64_t)nspans * S;
  const std::vector<float2> x = waveform(rng, (size_t)nchunks * kRxChunk + 8, omega, 38.f, 0.0013f, noise, false, shaped);
  orc_rx orx;
  orc_rx_init(&orx, t.oc, t.otrig.data(), 1);
  orc_rx_set_omega(&orx, omega);
  std::vector<uint32_t> osym((size_t)nchunks * kRxChunk);
  size_t ns = 0;
  orc_rx_run(&orx, reinterpret_cast<const float *>(x.data()), (size_t)nchunks * kRxChunk + 1, reinterpret_cast<uint8_t *>(osym.data()), &ns, nullptr, nullptr, nullptr, nullptr);

  RxState st0;
  memset(&st0, 0, sizeof st0);
  st0.est_insp = 75.0f * 75.0f; st0.agc_gain = 1;
  RxArgs a;
  memset(&a, 0, sizeof a);
  RxParams &p = a.p;
  p.cstln = reinterpret_cast<const CstlnCellDev *>(t.cst.cells.data());
  p.trig = reinterpret_cast<const float2 *>(t.trig.data());
  for (int k = 0; k < t.cst.nsymbols; ++k) { p.sym_re[k] = t.cst.sym_re[k]; p.sym_im[k] = t.cst.sym_im[k]; }
  p.nsymbols = t.cst.nsymbols; p.sampler = 1; p.omega = omega; p.min_freqw = orx.min_freqw; p.max_freqw = orx.max_freqw;
  p.freq_alpha = 0.04; p.freq_beta = 0.0012 / omega * 1.0f; p.gain_mu = 0.02 / (75.0f * 75.0f) * 2; p.kest = 0.01f;
  p.meas_decimation = 1048576; p.pe16 = t.pe.data(); p.slicer = 1;
  a.x = x.data();
  const size_t pe_bytes = (size_t)256 * kPeFoldPitch * 2;
  // ---- settling pass: the serial lane over the first chunks
  std::vector<uint32_t> head_sym((size_t)settle * kRxChunk + 4);
  RxSpanInfo hinfo; RxState settled;
  {
    RxArgs b = a;
    b.nchunks = settle; b.avail_chunks = nchunks; b.first_exact = 1; b.state_in = &st0; b.warm_in = &st0;
    b.span_chunks = settle; b.nspans = 1; b.span_cap = (uint32_t)(((size_t)settle * kRxChunk + 3) & ~(size_t)3);
    b.sym_out = head_sym.data(); b.info = &hinfo; b.state_end = &settled;
    std::vector<unsigned char> smem(pe_bytes + 2 * (kRxChunk + 8) * 8 + 256);
    emu::g_dyn_smem = smem.data();
    emu::launch(1, 128, [&] { dev::k_rx_serial<1>(b); });
  }
  size_t bad = 0;
  for (uint32_t i = 0; i < hinfo.n_out; ++i) bad += (head_sym[i] & 0xffffffu) != (osym[i] & 0xffffffu);
  CHECK(bad == 0, "%s: settling pass: %zu softsymbols differ", name, bad);
  // ---- the spans
  const uint32_t cap = ((uint32_t)((S + kRxVerifyChunks + 1) * kRxChunk / omega) + 64 + 3u) & ~3u;
  std::vector<uint32_t> out((size_t)nspans * cap, 0xdeadbeefu);
  std::vector<RxSpanInfo> info(nspans);
  std::vector<RxState> end(nspans), begin(nspans);
  std::vector<RxSeamSym> hlog((size_t)nspans * kRxSeamLog), tlog((size_t)nspans * kRxSeamLog);
  RxArgs b = a;
  b.nchunks = nchunks; b.avail_chunks = nchunks; b.chunk0 = settle; b.first_exact = 1; b.state_in = &settled; b.warm_in = &settled;
  b.state_chunk = settle; b.span_chunks = S; b.warm_chunks = W; b.nspans = nspans; b.span_cap = cap;
  b.sym_out = out.data(); b.info = info.data(); b.state_end = end.data(); b.state_begin = begin.data();
  b.head_log = hlog.data(); b.tail_log = tlog.data();
  const int warps = 2;
  std::vector<unsigned char> smem(pe_bytes + (size_t)warps * dev::kStages * 32 * (8 + 2) * 8 + 256);
  emu::g_dyn_smem = smem.data();
  emu::launch((nspans + warps * 32 - 1) / (warps * 32), warps * 32, [&] { dev::k_rx<8, 1>(b, nullptr, 0); });
  std::vector<RxSeam> seams(nspans);
  RxStitchArgs sa;
  memset(&sa, 0, sizeof sa);
  sa.info = info.data(); sa.head_log = hlog.data(); sa.tail_log = tlog.data(); sa.nspans = nspans;
  sa.nrot = t.cst.nrotations; sa.nsymbols = t.cst.nsymbols; sa.rot_perm = t.rot_perm.data(); sa.omega = omega;
  sa.seams = seams.data(); sa.strict = 1; sa.state_begin = begin.data(); sa.state_end = end.data();
  sa.tol_phase = 65536.0f / (float)sa.nrot / 16.0f; sa.tol_freqw = (p.max_freqw - p.min_freqw) / 64.0f;
  emu::launch((nspans - 1 + 3) / 4, 128, [&] { dev::k_rx_stitch(sa, nullptr, 0); });
  uint32_t failed = 0, rotated = 0, shifted = 0;
  for (uint32_t j = 0; j + 1 < nspans; ++j) { failed += !seams[j].ok; rotated += seams[j].rot != 0; shifted += seams[j].extend_prev || seams[j].skip_next; }
  // repair in stream order: the later span of a failed seam is re-run from its predecessor's end state, its next seam judged again
  uint32_t repaired = 0;
  for (uint32_t j = 0; j + 1 < nspans; ++j) {
    if (seams[j].ok) continue;
    const uint32_t list = j + 1;
    emu::launch(1, warps * 32, [&] { dev::k_rx<8, 1>(b, &list, 1); });
    ++repaired;
    const uint32_t sl[2] = {j, j + 1};
    emu::launch(1, 128, [&] { dev::k_rx_stitch(sa, sl, j + 2 < nspans ? 2 : 1); });
    CHECK(seams[j].ok && seams[j].mismatches == 0 && seams[j].rot == 0, "%s: seam %u does not verify after the exact re-run", name, j);
  }
  const unsigned tile = 64;
  const uint32_t nblk = (nspans + tile - 1) / tile;
  std::vector<uint64_t> span_offset((size_t)nspans + 1 + 2 * nblk + 4, 0);
  std::vector<uint32_t> span_skip(nspans, 0);
  std::vector<uint8_t> span_rot(nspans, 0);
  unsigned long long result[9] = {0};
  unsigned long long *totals = reinterpret_cast<unsigned long long *>(span_offset.data() + nspans + 1);
  emu::launch(nblk, tile, [&] { dev::k_rx_plan_local(info.data(), seams.data(), nspans, cap, sa.nrot, 0, 0, span_offset.data(), span_skip.data(), span_rot.data(), totals, result); });
  emu::launch(nblk, tile, [&] { dev::k_rx_plan_apply(nspans, sa.nrot, span_offset.data(), span_rot.data(), totals, result); });
  CHECK(result[0] == 0 && result[3] == 0, "%s: plan: %llu failed seams, %llu overflows", name, result[0], result[3]);
  CHECK(hinfo.n_out + result[1] == ns, "%s: %u + %llu symbols, oracle %zu", name, hinfo.n_out, result[1], ns);
  std::vector<uint32_t> flat((size_t)result[1] + 8, 0xdeadbeefu);
  RxCompactArgs ca;
  memset(&ca, 0, sizeof ca);
  ca.sym_in = out.data(); ca.span_cap = cap; ca.nspans = nspans; ca.span_offset = span_offset.data(); ca.span_skip = span_skip.data();
  ca.span_rot = span_rot.data(); ca.rot_perm = t.rot_perm.data(); ca.nsymbols = t.cst.nsymbols; ca.sym_out = flat.data();
  emu::launch(nspans, 64, [&] { dev::k_rx_compact(ca, result[1]); });
  size_t hard = 0, cost = 0;
  if (hinfo.n_out + result[1] == ns)
    for (size_t i = 0; i < result[1]; ++i) {
      const uint32_t g = flat[i], w = osym[hinfo.n_out + i];
      cost += (g & 0xffffu) != (w & 0xffffu);
      if (((g >> 16) & 0xffu) != ((w >> 16) & 0xffu)) {
        // a decision may only flip where it is marginal: the oracle's own confidence (cost, the gap between the two nearest
        // points: up to 2 * 53 * 53 * 2 for a sample on a constellation point) is a few per cent of the scale
        ++hard;
        const int oc = (int)(short)(w & 0xffffu), gc = (int)(short)(g & 0xffffu);
        CHECK(abs(oc) < 600 && abs(gc) < 600, "%s: symbol %zu flipped with costs %d (oracle) and %d", name, i, oc, gc);
      }
    }
  CHECK(hard <= max_flips, "%s: %zu of %llu hard decisions differ from the oracle", name, hard, result[1]);
  fprintf(stderr, "%s: %u spans of %u chunks after %u settling chunks: %u seams failed the strict rule and were repaired (%u rotated, %u shifted by a symbol), "
          "%llu symbols, hard decisions differing from the oracle: %zu, soft costs differing: %zu, equal so far: %s\n",
          name, nspans, S, settle, failed, rotated, shifted, result[1], hard, cost, g_fail ? "NO" : "yes");
  (void)repaired;
}

int main(int argc, char **argv) {
  const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  const bool quick = argc > 2;   // (under ThreadSanitizer: the bench configuration and the generic slicer once each)
  std::mt19937_64 rng(seed);
  const Tables qpsk = make_tables(1);
  run_case<1>(rng, qpsk, 1, 1.2f, quick ? 33 : 40, quick ? 2 : 3, quick ? 1 : 2, "QPSK, linear sampler, 1.2 samples per symbol, arithmetic slicer (the bench configuration)");
  run_case<0>(rng, qpsk, 1, 1.2f, quick ? 5 : 12, 4, 1, "QPSK, linear sampler, cell-table slicer");
  if (!quick) {
    run_case<1>(rng, qpsk, 0, 4.0f, 9, 5, 2, "QPSK, nearest sampler, 4 samples per symbol");
    run_case<1>(rng, qpsk, 2, 2.0f, 6, 6, 2, "QPSK, RRC sampler, 2 samples per symbol");
    const Tables psk8 = make_tables(2);
    run_case<0>(rng, psk8, 1, 2.0f, 7, 4, 2, "8PSK, linear sampler, 2 samples per symbol");
    const Tables apsk = make_tables(3);
    run_case<0>(rng, apsk, 2, 2.0f, 5, 4, 0, "16APSK, RRC sampler, no warm-up");
    run_case_hs(rng, 1.2f, 9, 4, 2, "--hs: fast_qpsk_receiver, 1.2 samples per symbol");
    run_fast_case(rng, qpsk, 2.5f, true, 0, 600, 64, 4, "FAST mode, QPSK at 1.2 samples per symbol, clean decisions");
    run_fast_case(rng, qpsk, 2.5f, false, 8, 600, 64, 4, "FAST mode, QPSK at 1.2 samples per symbol, heavy inter-symbol interference");
  }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_rx seed %llu: equal\n", (unsigned long long)seed);
  return 0;
}
