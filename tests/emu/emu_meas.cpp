// emu_meas.cpp -- TEST INFRASTRUCTURE.  Runs the telemetry kernels of leansdr_b200/csrc/k_spectrum.cu on the host
// (cuda_emu.h): k_meas_power (the reference's 4096- / 1024-point FFT in its butterfly order over shared memory, one
// barrier per stage, then |x / n|^2) and k_meas_ema (the running average of cnr_fft / spectrum and the band sums of
// do_cnr), against the oracle's cfft_engine + cnr_fft / spectrum (oracle/dvbs_oracle.c), float for float -- this file
// is built with -ffp-contract=off and the shim's fmul / fadd are single IEEE operations like the device intrinsics.
// The device text is the anonymous namespace of k_spectrum.cu (MEAS_DEV_INC).  Built with -fsanitize=thread the same
// run is the race check of the barrier-staged FFT.  Usage: emu_meas <seed>; exit code 0 = equal.
#include "cuda_emu.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../leansdr_b200/csrc/kernels.h"
extern "C" {
#include "../../oracle/dvbs_oracle.h"
}

inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline float __fdiv_rn(float a, float b) { return a / b; }
namespace ldvb {
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
namespace dev {
#include MEAS_DEV_INC
}
}  // namespace ldvb
using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)
static bool same_bits(float a, float b) { return memcmp(&a, &b, 4) == 0; }

int main(int argc, char **argv) {
  const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  std::mt19937_64 rng(seed);
  // omega_rev of the 4096-point engine, as ldvb_create builds it (dsp.h:70-76)
  std::vector<float2> tw(4096);
  for (int i = 0; i < 4096; ++i) { const float a = (float)(2.0 * M_PI * i / 4096); tw[i] = make_float2(cosf(a), -sinf(a)); }
  for (int logn : {12, 10}) {
    const int n = 1 << logn, npoints = 3;
    const float kavg = logn == 12 ? 0.1f : 0.5f;
    const float bandwidth = logn == 12 ? 2000e3f / 2400e3f : 0.f;     // leandvb.cc:327, 342
    // a carrier, a tone and noise
    std::vector<float> x(2 * (size_t)n * npoints);
    for (size_t i = 0; i < (size_t)n * npoints; ++i) {
      const float ph = 0.37f * (float)i, ph2 = 1.91f * (float)i;
      x[2 * i] = 40.f * cosf(ph) + 9.f * cosf(ph2) + (float)((int)(rng() % 2001) - 1000) * 0.01f;
      x[2 * i + 1] = 40.f * sinf(ph) + 9.f * sinf(ph2) + (float)((int)(rng() % 2001) - 1000) * 0.01f;
    }
    std::vector<uint64_t> start(npoints);
    for (int p = 0; p < npoints; ++p) start[p] = (uint64_t)p * n;
    std::vector<float> power((size_t)npoints * n, -1.f);
    MeasArgs a{};
    a.src.carry = nullptr; a.src.carry_count = 0;
    a.src.rest.head = x.data(); a.src.rest.head_count = (uint64_t)n * npoints; a.src.rest.main = nullptr; a.src.rest.c0 = 0;
    a.src.rest_off = 0; a.src.fmt = 5; a.src.scale = 1.f; a.src.rot_lut = nullptr; a.src.rot_index0 = 0;
    a.point_start = start.data(); a.npoints = npoints; a.logn = logn; a.twiddle_rev = tw.data(); a.power = power.data();
    emu::launch(npoints, 1024, [&] { dev::k_meas_power(a); });
    // running average + band sums
    std::vector<float> avg(n, 0.f), rows((size_t)npoints * n, -1.f), sums(3 * npoints, -1.f);
    int have = 0;
    MeasEmaArgs e{};
    e.power = power.data(); e.npoints = npoints; e.n = n; e.kavg = kavg; e.avg = avg.data(); e.have = &have;
    const float center_freq = 0.0213f;
    e.icf = (int)floor(center_freq * n + 0.5); e.bwslots = bandwidth > 0 ? (int)((bandwidth / 4) * n) : 0;
    e.sums = sums.data(); e.rows = rows.data();
    emu::launch(1, 1024, [&] { dev::k_meas_ema(e); });
    // the oracle, one block at a time
    orc_meas om;
    orc_meas_init(&om, n, bandwidth, kavg, n);
    for (int p = 0; p < npoints; ++p) {
      std::vector<float> blk(x.begin() + 2 * (size_t)n * p, x.begin() + 2 * (size_t)n * (p + 1));
      orc_fft_inplace(n, blk.data(), 1);
      size_t bad = 0;
      for (int i = 0; i < n; ++i) bad += !same_bits(blk[2 * i] * blk[2 * i] + blk[2 * i + 1] * blk[2 * i + 1], power[(size_t)p * n + i]);
      CHECK(bad == 0, "n %d point %d: %zu power bins differ from the oracle's FFT", n, p, bad);
      std::vector<float> out(n + 4);
      size_t consumed = 0;
      const size_t got = orc_meas_run(&om, x.data() + 2 * (size_t)n * p, n, center_freq, out.data(), 1, &consumed);
      CHECK(got == 1 && consumed == (size_t)n, "n %d point %d: oracle measured %zu points", n, p, got);
      bad = 0;
      for (int i = 0; i < n; ++i) bad += !same_bits(om.avgpower[i], rows[(size_t)p * n + i]);
      CHECK(bad == 0, "n %d point %d: %zu bins of the running average differ", n, p, bad);
      if (e.bwslots) {   // do_cnr (sdr.h:1306-1331): the three band averages, then the oracle's dB value from them
        const float c2n2 = sums[3 * p], nl = sums[3 * p + 1], nr = sums[3 * p + 2];
        const float n2 = (nl + nr) / 2, c2 = c2n2 - n2;
        const float cnr = (c2 > 0 && n2 > 0) ? 10 * logf(c2 / n2) / logf(10) : -50;
        CHECK(same_bits(cnr, out[0]), "point %d: cnr %g vs oracle %g", p, cnr, out[0]);
      }
    }
    CHECK(have == 1, "have flag");
  }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_meas seed %llu: equal\n", (unsigned long long)seed);
  return 0;
}
