// emu_notch_detect.cpp -- TEST INFRASTRUCTURE.  Runs k_notch_detect (leansdr_b200/csrc/k_notch.cu: auto_notch::detect,
// sdr.h:76-118 -- the 4096-point FFT of cfft_engine in shared memory, glibc's hypotf in double, the nslots largest bins
// with their neighbours blanked) on the host (cuda_emu.h) against the oracle's auto_notch, whose detect() is made to
// run on every block (decimation = 4096).  The device text is the part of k_notch.cu from load_sample to the end of
// k_notch_detect (DETECT_DEV_INC).  Built with -fsanitize=thread the same run is the kernel's race check.
// Usage: emu_notch_detect <seed>; exit code 0 = equal.
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
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline float __double2float_rn(double a) { return (float)a; }
namespace ldvb {
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
namespace dev {
#include DETECT_DEV_INC
}
}  // namespace ldvb
using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

int main(int argc, char **argv) {
  const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  std::mt19937_64 rng(seed);
  std::vector<float2> tw(4096);
  for (int i = 0; i < 4096; ++i) { const float a = (float)(2.0 * M_PI * i / 4096); tw[i] = make_float2(cosf(a), -sinf(a)); }
  for (int nslots = 1; nslots <= 4; ++nslots) {
    for (int fmt : {5, 0}) {
      const int nblk = 3;
      // tones of different strengths on random bins (some adjacent to each other: the blanking rule), noise
      std::vector<float> x(2 * 4096 * (size_t)nblk);
      std::vector<unsigned char> u8(2 * 4096 * (size_t)nblk);
      for (int b = 0; b < nblk; ++b) {
        int bins[5]; float amp[5];
        for (int t = 0; t < 5; ++t) { bins[t] = (int)(rng() % 4096); amp[t] = 6.f + (float)(rng() % 30); }
        bins[1] = (bins[0] + 1) & 4095;                                      // a neighbour of the strongest candidate
        for (int i = 0; i < 4096; ++i) {
          float re = (float)((int)(rng() % 2001) - 1000) * 0.004f, im = (float)((int)(rng() % 2001) - 1000) * 0.004f;
          for (int t = 0; t < 5; ++t) { const float ph = (float)(2.0 * M_PI * bins[t] * i / 4096.0); re += amp[t] * cosf(ph); im += amp[t] * sinf(ph); }
          const size_t k = (size_t)b * 4096 + i;
          if (fmt == 0) {
            u8[2 * k] = (unsigned char)std::min(255.f, std::max(0.f, 128.f + re)); u8[2 * k + 1] = (unsigned char)std::min(255.f, std::max(0.f, 128.f + im));
          } else { x[2 * k] = re; x[2 * k + 1] = im; }
        }
      }
      if (fmt == 0) orc_cconvert(u8.data(), 0, x.data(), 4096 * (size_t)nblk);
      std::vector<uint64_t> blocks(nblk);
      for (int b = 0; b < nblk; ++b) blocks[b] = (uint64_t)b;
      std::vector<int32_t> bins_out(nblk * nslots, -1);
      NotchDetectArgs a{};
      a.src.head = fmt == 0 ? (const void *)u8.data() : (const void *)x.data(); a.src.head_count = 4096 * (uint64_t)nblk; a.src.main = nullptr; a.src.c0 = 0;
      a.fmt = fmt; a.scale = 1.f; a.block_index = blocks.data(); a.ndetect = nblk; a.nslots = nslots; a.twiddle_rev = tw.data(); a.bins_out = bins_out.data();
      emu::launch(nblk, 1024, [&] { dev::k_notch_detect(a); });
      orc_notch on;
      orc_notch_init(&on, nslots);
      on.decimation = 4096;                                                  // detect() on every block
      std::vector<float> y(2 * 4096);
      for (int b = 0; b < nblk; ++b) {
        orc_notch_run(&on, x.data() + 2 * 4096 * (size_t)b, 4096, y.data());
        for (int s = 0; s < nslots; ++s)
          CHECK(bins_out[b * nslots + s] == on.slots[s].i, "nslots %d fmt %d block %d slot %d: bin %d vs oracle %d", nslots, fmt, b, s, bins_out[b * nslots + s], on.slots[s].i);
      }
    }
  }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_notch_detect seed %llu: equal\n", (unsigned long long)seed);
  return 0;
}
