// emu_notchfir.cpp -- TEST INFRASTRUCTURE.  Runs the DEFAULT notch kernel of the receive chain -- k_notch_fir (leansdr_b200/
// csrc/k_notchfir.cu: auto_notch::process and the fir_filter behind it fused, one chain warp + eight worker warps per
// 16 segments, a block barrier per 64-sample tile, five stages of asynchronous row copies) with k_fir_edges, behind
// k_notch_guess and in front of k_notch_verify (k_notch.cu) -- on the host (cuda_emu.h) against the oracle's auto_notch
// followed by its fir_filter, FLOAT FOR FLOAT (-ffp-contract=off; asynchronous copies performed at issue time).
//   * two consecutive batches: the first with no carried samples, the second with the fir_n notched samples and the notch
//     estimates the first one left (carry_out, exit state);
//   * segments of one block with two warm-up blocks (all merge), and without warm-up (none merges: every segment is
//     re-run from its predecessor's exit state in rounds, rewriting its edge samples);
//   * 5 real taps (the bench configuration), 13 retuned taps, no FIR (plain notch through the same kernel), 1 and 2 slots;
//   * the blocks cnr_fft / spectrum will read are dumped on the side.
// Device text: the anonymous namespaces of k_notch.cu (NOTCH_DEV_INC) and k_notchfir.cu (NOTCHFIR_DEV_INC) up to
// their templated launchers, each behind notch_common.cuh.  Built with -fsanitize=thread the same run is the race check
// of the warp-specialised kernel.  Usage: emu_notchfir <seed> [quick]; exit code 0 = equal.
#include "cuda_emu.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

extern "C" {
#include "../../oracle/dvbs_oracle.h"
}

inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline float __double2float_rn(double a) { return (float)a; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
template <class T> inline T __ldcs(const T *p) { return *p; }
#include "../../leansdr_b200/csrc/kernels.h"
#include "../../leansdr_b200/csrc/tables.h"
namespace ldvb {
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
inline float2 cmul(float2 a, float2 b) { return make_float2(fsub(fmul(a.x, b.x), fmul(a.y, b.y)), fadd(fmul(a.x, b.y), fmul(a.y, b.x))); }
inline void st_stream(float2 *p, float2 v) { *p = v; }
inline void st_stream(float4 *p, float4 v) { *p = v; }
inline void cp_async16(void *dst, const void *src) { memcpy(dst, src, 16); }
inline void cp_async16_ca(void *dst, const void *src) { memcpy(dst, src, 16); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
namespace dev {
#include NOTCH_DEV_INC
}
namespace devf {
#include NOTCHFIR_DEV_INC
}
}  // namespace ldvb
using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

template <int NSLOTS, bool FIR>
static void run_case(std::mt19937_64 &rng, int fir_n, float retune, uint32_t warm_blocks) {
  const uint64_t nb[2] = {19, 7};                      // blocks per batch (19 segments: two CTAs, the second one ragged)
  int bins[4];
  for (int s = 0; s < NSLOTS; ++s) bins[s] = (int)(rng() % 4096);
  std::vector<float2> tables((size_t)(NSLOTS + 1) * 4096, make_float2(0.f, 0.f));
  orc_notch on;
  orc_notch_init(&on, NSLOTS);
  for (int s = 0; s < NSLOTS; ++s) {
    on.slots[s].i = bins[s];
    for (int i = 0; i < 4096; ++i) {
      const float ang = (float)(2 * M_PI * bins[s] * i / 4096);
      on.slots[s].expj[2 * i] = cosf(ang); on.slots[s].expj[2 * i + 1] = sinf(ang);
      tables[(size_t)(s + 1) * 4096 + i] = make_float2(cosf(ang), sinf(ang));
    }
  }
  // fir_filter behind the notch (decimation 1)
  std::vector<float> coeffs(fir_n > 0 ? fir_n : 1), taps;
  orc_fir of{};
  if (FIR) {
    for (int i = 0; i < fir_n; ++i) coeffs[i] = (float)((int)(rng() % 2001) - 1000) * 0.0009f;
    orc_fir_init(&of, (unsigned)fir_n, coeffs.data(), 1);
    orc_fir_set_freq(&of, retune);
    taps = shift_taps(coeffs, retune);
    CHECK(memcmp(taps.data(), of.shifted, 8 * (size_t)fir_n) == 0, "shifted taps");
  }
  bool real_taps = true;
  for (int i = 0; i < fir_n; ++i) real_taps = real_taps && taps[2 * i + 1] == 0.0f;
  std::vector<float> weights(8192);
  { const double c1 = (double)(1.0f - 0.002f); for (int m = 0; m < 8192; ++m) weights[m] = (float)pow(c1, (double)m); }

  NotchState st{};
  st.phase = 0; st.gain = 1.0f;
  for (int s = 0; s < kNotchMaxSlots; ++s) { st.slot[s].bin = s < NSLOTS ? bins[s] : -1; st.slot[s].est_re = 0; st.slot[s].est_im = 0; }
  std::vector<float2> carry(kFirFuseMaxTaps, make_float2(0.f, 0.f));
  std::vector<float> fir_hist;                         // the oracle's unread notched samples (fir_n of them after a batch)
  uint32_t ncarry = 0, repaired_total = 0;
  for (int batch = 0; batch < 2; ++batch) {
    const uint64_t nblocks = nb[batch];
    const size_t n = 4096 * (size_t)nblocks;
    std::vector<float> x(2 * n);
    for (size_t i = 0; i < n; ++i) {
      float re = (float)((int)(rng() % 2001) - 1000) * 0.01f, im = (float)((int)(rng() % 2001) - 1000) * 0.01f;
      for (int s = 0; s < NSLOTS; ++s) { const float ph = (float)(2.0 * M_PI * bins[s] * (double)(i % 4096) / 4096.0) + 0.3f * s; re += (30.f + 5.f * s) * cosf(ph); im += (30.f + 5.f * s) * sinf(ph); }
      x[2 * i] = re; x[2 * i + 1] = im;
    }
    // ---- the oracle: notch, then the FIR over [unread notched samples | this batch]
    std::vector<float> notched(2 * n);
    orc_notch_run(&on, x.data(), n, notched.data());
    std::vector<float> want;
    if (FIR) {
      std::vector<float> u(fir_hist);
      u.insert(u.end(), notched.begin(), notched.end());
      want.resize(u.size() + 16);
      size_t consumed = 0;
      const size_t got = orc_fir_run(&of, u.data(), u.size() / 2, want.data(), &consumed);
      want.resize(2 * got);
      fir_hist.assign(u.begin() + 2 * consumed, u.end());
      CHECK(fir_hist.size() == 2 * (size_t)fir_n, "oracle fir_filter left %zu samples unread", fir_hist.size() / 2);
    } else {
      want = notched;
    }
    // ---- the kernels
    NotchEpoch ep{};
    ep.first_block = 0;
    for (int s = 0; s < NSLOTS; ++s) { ep.bin[s] = bins[s]; ep.reset[s] = 0; ep.table_index[s] = (uint32_t)(s + 1); }
    NotchFirArgs fa{};
    NotchApplyArgs &a = fa.n;
    std::vector<float> out(2 * n + 8, -7.f), y(2 * (n + kFirFuseMaxTaps) + 8, -7.f);
    a.src.head = x.data(); a.src.head_count = n; a.src.main = nullptr; a.src.c0 = 0; a.fmt = 5; a.scale = 1.f;
    a.out = reinterpret_cast<float2 *>(out.data()); a.nblocks = nblocks; a.nslots = NSLOTS; a.k = 0.002f; a.gain = 1.0f;
    a.w_block = (float)pow((double)(1.0f - 0.002f), 4096.0);
    a.expj_tables = tables.data(); a.epochs = &ep; a.nepochs = 1; a.block0 = 0; a.first_exact = 1;
    a.seg_blocks = 1; a.warm_blocks = warm_blocks; a.nsegs = (uint32_t)nblocks; a.state_in = &st;
    std::vector<float2> entry((size_t)a.nsegs * kNotchMaxSlots + 8), exitv(entry.size());
    std::vector<uint8_t> exact(a.nsegs + 8, 9);
    a.seg_entry = entry.data(); a.seg_exit = exitv.data(); a.seg_exact = exact.data();
    std::vector<float2> guess((size_t)(nblocks + 1) * kNotchMaxSlots, make_float2(0.f, 0.f));
    std::vector<float2> edge((size_t)a.nsegs * kNotchEdge + 8, make_float2(-3.f, -3.f));
    const std::vector<uint64_t> dump_blocks = {1, nblocks - 2};
    std::vector<float2> dump(dump_blocks.size() * 4096, make_float2(-5.f, -5.f));
    fa.fir_n = FIR ? fir_n : 0; fa.real_taps = real_taps ? 1 : 0; fa.taps = reinterpret_cast<const float2 *>(taps.data());
    fa.y = reinterpret_cast<float2 *>(y.data()); fa.carry = ncarry; fa.carry_in = carry.data(); fa.carry_out = carry.data();
    fa.edge = edge.data(); fa.dump_blocks = dump_blocks.data(); fa.ndump = (int)dump_blocks.size(); fa.dump = dump.data();
    using SM = devf::FSmem<5, NSLOTS, FIR>;
    std::vector<unsigned char> dyn(SM::total + 256);
    emu::g_dyn_smem = reinterpret_cast<unsigned char *>(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
    {  // launch_notch_guess
      const uint64_t lead = (uint64_t)a.warm_blocks + 2;
      const uint64_t first = a.block0 > lead ? a.block0 - lead : 0;
      if (a.nblocks > first) emu::launch((unsigned)(a.nblocks - first), 128, [&] { dev::k_notch_guess<5, NSLOTS>(a, first, guess.data(), weights.data()); });
    }
    emu::launch((a.nsegs + devf::kFRows - 1) / devf::kFRows, devf::kFThreads, [&] { devf::k_notch_fir<5, NSLOTS, FIR>(fa, nullptr, 0, guess.data()); });
    uint32_t repaired = 0;                             // notch_verify_repair (pipeline.cu)
    for (int round = 0; round < 64 && a.nsegs > 1; ++round) {
      uint32_t nfail = 0;
      emu::launch((a.nsegs + 255) / 256, 256, [&] { dev::k_notch_verify(a.seg_entry, a.seg_exit, a.seg_exact, a.nsegs, a.nslots, &nfail); });
      if (!nfail) break;
      std::vector<uint32_t> todo;
      bool prev_failed = false;
      for (uint32_t j = 1; j < a.nsegs; ++j) {
        bool same = exact[j] != 0;
        if (!same) same = memcmp(&entry[(size_t)j * kNotchMaxSlots], &exitv[(size_t)(j - 1) * kNotchMaxSlots], 8 * (size_t)NSLOTS) == 0;
        if (!same && !prev_failed) todo.push_back(j);
        prev_failed = !same;
      }
      CHECK(!todo.empty(), "verification counted %u failures but none can be repaired", nfail);
      if (todo.empty()) break;
      const uint32_t nl = (uint32_t)todo.size();
      emu::launch((nl + devf::kFRows - 1) / devf::kFRows, devf::kFThreads, [&] { devf::k_notch_fir<5, NSLOTS, FIR>(fa, todo.data(), nl, nullptr); });
      repaired += nl;
      for (uint32_t j : todo) { memcpy(&entry[(size_t)j * kNotchMaxSlots], &exitv[(size_t)(j - 1) * kNotchMaxSlots], 8 * kNotchMaxSlots); exact[j] = 1; }
    }
    repaired_total += repaired;
    if (FIR) {  // launch_fir_edges
      const uint32_t threads = a.nsegs * (uint32_t)(fir_n > 1 ? fir_n - 1 : 1);
      emu::launch((threads + 255) / 256, 256, [&] { devf::k_fir_edges(fa); });
    }
    // ---- compare
    const std::vector<float> &got = FIR ? y : out;
    const size_t nout = want.size();
    if (FIR) CHECK(nout == 2 * ((size_t)ncarry + n - (size_t)fir_n), "batch %d: oracle produced %zu outputs", batch, nout / 2);
    size_t bad = 0, first_bad = 0;
    for (size_t i = 0; i < nout; ++i) if (memcmp(&got[i], &want[i], 4) != 0) { if (!bad) first_bad = i; ++bad; }
    CHECK(bad == 0, "%d slots, %d taps, warm-up %u, batch %d: %zu of %zu floats differ from the oracle (first at output %zu: %g vs %g), %u repaired",
          NSLOTS, FIR ? fir_n : 0, warm_blocks, batch, bad, nout, first_bad / 2, got[first_bad], want[first_bad], repaired);
    CHECK(got[nout] == -7.f, "batch %d: wrote past the end", batch);
    for (size_t d = 0; d < dump_blocks.size(); ++d)
      CHECK(memcmp(&dump[d * 4096], &notched[2 * 4096 * dump_blocks[d]], 8 * 4096) == 0, "batch %d: telemetry dump of block %llu", batch, (unsigned long long)dump_blocks[d]);
    for (int s = 0; s < NSLOTS; ++s) {
      const float2 e = exitv[(size_t)(a.nsegs - 1) * kNotchMaxSlots + s];
      CHECK(memcmp(&e.x, &on.slots[s].estim_re, 4) == 0 && memcmp(&e.y, &on.slots[s].estim_im, 4) == 0, "batch %d: carried estimate of slot %d", batch, s);
      st.slot[s].est_re = e.x; st.slot[s].est_im = e.y;                 // what run_notch carries to the next batch
    }
    if (FIR) {
      CHECK(memcmp(carry.data(), fir_hist.data(), 8 * (size_t)fir_n) == 0, "batch %d: carried notched samples", batch);
      ncarry = (uint32_t)fir_n;
    }
  }
  if (warm_blocks == 0) CHECK(repaired_total > 0, "no warm-up, yet nothing had to be repaired: the repair path did not run");
  fprintf(stderr, "  %d slot(s), %d taps%s, %u warm-up blocks: %u segments repaired\n", NSLOTS, FIR ? fir_n : 0, retune != 0.f ? " (retuned)" : "", warm_blocks, repaired_total);
}

int main(int argc, char **argv) {
  const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  const bool quick = argc > 2;
  std::mt19937_64 rng(seed);
  run_case<1, true>(rng, 5, 0.f, 2);          // the bench configuration
  run_case<1, true>(rng, 5, 0.f, 0);          // ... every segment repaired
  if (!quick) {
    run_case<2, true>(rng, 13, 0.027f, 2);    // two slots, retuned (complex) taps
    run_case<1, false>(rng, 0, 0.f, 2);       // plain notch through the same kernel
    run_case<2, false>(rng, 0, 0.f, 0);
  }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_notchfir seed %llu: equal\n", (unsigned long long)seed);
  return 0;
}
