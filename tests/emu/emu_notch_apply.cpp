// emu_notch_apply.cpp -- TEST INFRASTRUCTURE.  Runs the notch kernels of leansdr_b200/csrc/k_notch.cu on the host (cuda_emu.h)
// against the oracle's auto_notch::process (sdr.h:119-138), FLOAT FOR FLOAT: this file is built with -ffp-contract=off,
// the shim's fmul / fadd are single IEEE operations, and the asynchronous row copies are performed at issue time (one of
// the schedules the hardware may choose).
//   exact        one segment from the carried state over the whole batch: k_notch_apply's arithmetic and streaming;
//   speculative  one segment per block: start states guessed by k_notch_guess, two warm-up blocks, entry(j) compared
//                with exit(j-1) bit for bit (k_notch_verify), failed segments re-run from their predecessor's exit
//                state in rounds -- the scheme of run_notch / notch_verify_repair (pipeline.cu) -- and the result must
//                be the oracle's serial run whatever merged and whatever did not (without warm-up blocks nothing
//                merges: the repair path).
// The device text is the first anonymous namespace of k_notch.cu up to its templated launchers, plus k_notch_verify
// (NOTCH_DEV_INC), with notch_common.cuh in front of it.  Built with -fsanitize=thread the same run is the race check.
// Usage: emu_notch_apply <seed> [quick]; exit code 0 = equal.
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
}  // namespace ldvb
using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

template <int NSLOTS>
static void run_case(std::mt19937_64 &rng, bool speculative, uint32_t warm_blocks = 2) {
  const uint64_t nblocks = speculative ? 9 : 3;
  const size_t n = 4096 * (size_t)nblocks;
  // a strong carrier per slot on the slot's bin plus noise: what the notch is there to remove
  int bins[4];
  for (int s = 0; s < NSLOTS; ++s) bins[s] = (int)(rng() % 4096);
  std::vector<float> x(2 * n);
  for (size_t i = 0; i < n; ++i) {
    float re = (float)((int)(rng() % 2001) - 1000) * 0.01f, im = (float)((int)(rng() % 2001) - 1000) * 0.01f;
    for (int s = 0; s < NSLOTS; ++s) { const float ph = (float)(2.0 * M_PI * bins[s] * (double)(i % 4096) / 4096.0) + 0.3f * s; re += (30.f + 5.f * s) * cosf(ph); im += (30.f + 5.f * s) * sinf(ph); }
    x[2 * i] = re; x[2 * i + 1] = im;
  }
  // tables (sdr.h:104-108), the same for both sides; table 0 is the all-zero table of a slot that never detected
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
  std::vector<float> want(2 * n);
  const size_t used = orc_notch_run(&on, x.data(), n, want.data());
  CHECK(used == n, "oracle consumed %zu of %zu samples", used, n);

  NotchEpoch ep{};
  ep.first_block = 0;
  for (int s = 0; s < NSLOTS; ++s) { ep.bin[s] = bins[s]; ep.reset[s] = 0; ep.table_index[s] = (uint32_t)(s + 1); }
  NotchState st0{};
  st0.phase = 0; st0.gain = 1.0f;
  for (int s = 0; s < kNotchMaxSlots; ++s) { st0.slot[s].bin = s < NSLOTS ? bins[s] : -1; st0.slot[s].est_re = 0; st0.slot[s].est_im = 0; }
  std::vector<float> out(2 * n + 8, -7.f);
  NotchApplyArgs a{};
  a.src.head = x.data(); a.src.head_count = n; a.src.main = nullptr; a.src.c0 = 0; a.fmt = 5; a.scale = 1.f;
  a.out = reinterpret_cast<float2 *>(out.data()); a.nblocks = nblocks; a.nslots = NSLOTS; a.k = 0.002f; a.gain = 1.0f;
  a.w_block = (float)pow((double)(1.0f - 0.002f), 4096.0);
  a.expj_tables = tables.data(); a.epochs = &ep; a.nepochs = 1; a.block0 = 0; a.first_exact = 1;
  a.seg_blocks = speculative ? 1 : (uint32_t)nblocks; a.warm_blocks = warm_blocks;
  a.nsegs = (uint32_t)((nblocks + a.seg_blocks - 1) / a.seg_blocks);
  a.state_in = &st0;
  std::vector<float2> entry((size_t)a.nsegs * kNotchMaxSlots + 8), exitv(entry.size());
  std::vector<uint8_t> exact(a.nsegs + 8, 9);
  a.seg_entry = entry.data(); a.seg_exit = exitv.data(); a.seg_exact = exact.data();
  std::vector<float2> guess((size_t)(nblocks + 1) * kNotchMaxSlots, make_float2(0.f, 0.f));
  std::vector<float> weights(8192);
  { const double c1 = (double)(1.0f - 0.002f); for (int m = 0; m < 8192; ++m) weights[m] = (float)pow(c1, (double)m); }
  std::vector<unsigned char> dyn(dev::NotchSmem<NSLOTS>::total + 256);
  emu::g_dyn_smem = reinterpret_cast<unsigned char *>(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
  const unsigned per_block = dev::kNWarps * 32;
  {  // launch_notch_guess
    const uint64_t lead = (uint64_t)a.warm_blocks + 2;
    const uint64_t first = a.block0 > lead ? a.block0 - lead : 0;
    if (a.nblocks > first) emu::launch((unsigned)(a.nblocks - first), 128, [&] { dev::k_notch_guess<5, NSLOTS>(a, first, guess.data(), weights.data()); });
  }
  emu::launch((a.nsegs + per_block - 1) / per_block, per_block, [&] { dev::k_notch_apply<5, NSLOTS>(a, nullptr, 0, guess.data()); });
  // notch_verify_repair
  uint32_t repaired = 0;
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
    emu::launch((nl + per_block - 1) / per_block, per_block, [&] { dev::k_notch_apply<5, NSLOTS>(a, todo.data(), nl, nullptr); });
    repaired += nl;
    for (uint32_t j : todo) { memcpy(&entry[(size_t)j * kNotchMaxSlots], &exitv[(size_t)(j - 1) * kNotchMaxSlots], 8 * kNotchMaxSlots); exact[j] = 1; }
  }
  size_t bad = 0, first_bad = 0;
  for (size_t i = 0; i < 2 * n; ++i) if (memcmp(&out[i], &want[i], 4) != 0) { if (!bad) first_bad = i; ++bad; }
  CHECK(bad == 0, "%d slots, %s: %zu of %zu floats differ from the oracle (first at sample %zu: %g vs %g), %u segments repaired", NSLOTS,
        speculative ? "speculative" : "exact", bad, 2 * n, first_bad / 2, out[first_bad], want[first_bad], repaired);
  CHECK(out[2 * n] == -7.f, "wrote past the end");
  for (int s = 0; s < NSLOTS; ++s) {
    const float2 e = exitv[(size_t)(a.nsegs - 1) * kNotchMaxSlots + s];
    CHECK(memcmp(&e.x, &on.slots[s].estim_re, 4) == 0 && memcmp(&e.y, &on.slots[s].estim_im, 4) == 0, "%d slots: carried estimate of slot %d", NSLOTS, s);
  }
  if (speculative && warm_blocks == 0) CHECK(repaired > 0, "no warm-up, yet nothing had to be repaired: the repair path did not run");
  fprintf(stderr, "  %d slot(s), %s, %u warm-up blocks: %u of %u segments repaired\n", NSLOTS, speculative ? "speculative" : "exact", warm_blocks, repaired, a.nsegs);
}

int main(int argc, char **argv) {
  const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  std::mt19937_64 rng(seed);
  const bool quick = argc > 2;   // (under ThreadSanitizer: one exact, one merging and one repair case -- the barriers and
                                 //  exchanges do not depend on the slot count; the full list runs in the plain build)
  if (!quick) {
    run_case<1>(rng, false);
    run_case<1>(rng, true);
  }
  run_case<2>(rng, false);
  run_case<2>(rng, true);
  if (!quick) run_case<3>(rng, true);
  run_case<1>(rng, true, 0);   // the guess alone does not merge: every segment is re-run exactly, in rounds
  if (!quick) run_case<2>(rng, true, 0);
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_notch_apply seed %llu: equal\n", (unsigned long long)seed);
  return 0;
}
