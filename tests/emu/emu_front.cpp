// emu_front.cpp -- TEST INFRASTRUCTURE.  Runs k_frontend (leansdr_b200/csrc/k_frontend.cu: cconverter / scaler -> rotator ->
// fir_filter with decimation, fused, the tile's raw span brought in by one bulk copy signalled on an mbarrier) on the
// host (cuda_emu.h) against the oracle's chain of the same runnables, FLOAT FOR FLOAT: this file is built with
// -ffp-contract=off and the shim's fmul / fadd / cmul are single IEEE operations like the device's _rn intrinsics.  The
// bulk copy is performed at issue time by the issuing thread and the mbarrier is a release / acquire flag -- one of the
// schedules the hardware may choose.  The device text is the anonymous namespace of k_frontend.cu (FRONT_DEV_INC, its
// dynamic shared memory pointed at the shim's buffer); the launch geometry repeats launch_frontend.  Built with
// -fsanitize=thread the same run is the kernel's race check.  Usage: emu_front <seed>; exit code 0 = equal.
#include "cuda_emu.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../leansdr_b200/csrc/kernels.h"
#include "../../leansdr_b200/csrc/tables.h"
extern "C" {
#include "../../oracle/dvbs_oracle.h"
}

template <class T> inline void __stcs(T *p, T v) { *p = v; }
namespace ldvb {
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
inline float2 cmul(float2 a, float2 b) { return make_float2(fsub(fmul(a.x, b.x), fmul(a.y, b.y)), fadd(fmul(a.x, b.y), fmul(a.y, b.x))); }
inline void st_stream(float2 *p, float2 v) { *p = v; }
inline void st_stream(float4 *p, float4 v) { *p = v; }
// mbarrier + 1-D bulk copy (common.cuh): copy at issue, then publish; waiters acquire
inline void mbar_init(uint64_t *bar, uint32_t) { __atomic_store_n(bar, (uint64_t)0, __ATOMIC_RELEASE); }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t *, uint32_t) {}
inline void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  memcpy(dst, src, bytes);
  __atomic_store_n(bar, (uint64_t)1, __ATOMIC_RELEASE);
}
inline void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (__atomic_load_n(bar, __ATOMIC_ACQUIRE) == (uint64_t)parity) std::this_thread::yield();
}
namespace dev {
#include FRONT_DEV_INC
}
}  // namespace ldvb
using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

struct Case { int fmt; float scale; float rot; unsigned ntaps; unsigned decim; float retune; const char *name; };

int main(int argc, char **argv) {
  const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  std::mt19937_64 rng(seed);
  const Case cases[] = {
      {4, 1.0f, 0.f, 5, 1, 0.f, "f32, 5 real taps (the bench configuration)"},
      {0, 1.0f, 0.0123f, 13, 2, 0.031f, "u8, rotator, 13 retuned taps, decimation 2"},
      {3, 1.0f, 0.f, 0, 3, 0.f, "s16, decimator alone"},
      {4, 0.5f, 0.f, 31, 1, -0.07f, "f32 scaled, 31 retuned taps"},
      {2, 1.0f, -0.2f, 7, 1, 0.f, "u16, rotator, 7 real taps"},
      {1, 1.0f, 0.f, 40, 5, 0.f, "s8, 40 taps, decimation 5"},
  };
  for (const Case &c : cases) {
    const unsigned N = c.ntaps, D = c.decim;
    const uint64_t count = 4096 * 2 + 700 + rng() % 900;               // outputs: two full tiles and a ragged one
    const uint64_t nin = count * D + N + 64;
    // raw input in the format of the case, and its cf32 image through the oracle's cconverter / scaler / rotator
    std::vector<float> x(2 * nin);
    std::vector<unsigned char> raw;
    if (c.fmt == 4) {
      raw.resize(8 * nin);
      float *f = reinterpret_cast<float *>(raw.data());
      for (uint64_t i = 0; i < 2 * nin; ++i) f[i] = (float)((int)(rng() % 20001) - 10000) * 0.013f;
      orc_scale(f, c.scale, x.data(), nin);
    } else {
      const int bytes = (c.fmt <= 1) ? 1 : 2;
      raw.resize((size_t)2 * bytes * nin);
      for (auto &b : raw) b = (unsigned char)rng();
      orc_cconvert(raw.data(), c.fmt, x.data(), nin);
    }
    std::vector<float> rot_lut;
    static orc_rotator orot;
    if (c.rot != 0.f) {
      rot_lut = make_rotator_lut(c.rot);
      orc_rotator_init(&orot, c.rot);
      CHECK(memcmp(rot_lut.data(), orot.lut_cos, 65536 * 4) == 0 && memcmp(rot_lut.data() + 65536, orot.lut_sin, 65536 * 4) == 0, "%s: rotator table", c.name);
      std::vector<float> y(2 * nin);
      orc_rotator_run(&orot, x.data(), y.data(), nin);
      x.swap(y);
    }
    // taps: an arbitrary low-pass-like vector, shifted like fir_filter::set_freq (dsp.h:270-280)
    std::vector<float> coeffs(N);
    for (unsigned i = 0; i < N; ++i) coeffs[i] = (float)((int)(rng() % 2001) - 1000) * 0.0007f;
    std::vector<float> want(2 * (count + 8), 0.f);
    std::vector<float> taps;
    if (N) {
      orc_fir f;
      orc_fir_init(&f, N, coeffs.data(), D);
      orc_fir_set_freq(&f, c.retune);
      taps = shift_taps(coeffs, c.retune);
      CHECK(taps.size() == 2 * N && memcmp(taps.data(), f.shifted, 8 * N) == 0, "%s: shifted taps", c.name);
      size_t consumed = 0;
      std::vector<float> all(2 * (nin / D + 8));
      const size_t got = orc_fir_run(&f, x.data(), nin, all.data(), &consumed);
      CHECK(got >= count, "%s: oracle produced %zu outputs", c.name, got);
      memcpy(want.data(), all.data(), 8 * count);
    } else {
      size_t consumed = 0;
      std::vector<float> all(2 * (nin / D + 8));
      const size_t got = orc_decimate(x.data(), nin, D, all.data(), &consumed);
      CHECK(got >= count, "%s: oracle decimated %zu outputs", c.name, got);
      memcpy(want.data(), all.data(), 8 * count);
    }
    bool real_taps = true;
    for (unsigned i = 0; i < N; ++i) real_taps = real_taps && taps[2 * i + 1] == 0.0f;
    // the kernel (launch_frontend's geometry)
    std::vector<float> out(2 * (count + 8), -7.f);
    FrontendArgs a{};
    a.src.head = raw.data(); a.src.head_count = nin; a.src.main = nullptr; a.src.c0 = 0;
    a.fmt = c.fmt; a.scale = c.scale; a.rot_lut = c.rot != 0.f ? rot_lut.data() : nullptr; a.rot_index0 = 0;
    a.taps = reinterpret_cast<const float2 *>(taps.data()); a.ntaps = N; a.decim = D; a.real_taps = real_taps ? 1 : 0;
    a.out = reinterpret_cast<float2 *>(out.data()); a.count = count;
    a.bytes_per_sample = (c.fmt <= 1) ? 2 : (c.fmt <= 3 ? 4 : 8);
    uint32_t tile = 256 * 16;
    while (tile > 256 && (uint64_t)(tile - 1) * D + N + 16 > 8192) tile -= 256;
    a.tile_out = tile;
    const uint32_t span_max = (tile - 1) * D + (N ? N : 1) + 16;
    a.max_raw_bytes = (span_max * a.bytes_per_sample + 15u) & ~15u;
    size_t smem = (((size_t)N * 8 + 127) & ~(size_t)127) + (((size_t)a.max_raw_bytes + 127) & ~(size_t)127);
    if (c.fmt < 4) smem += (size_t)span_max * 8;
    std::vector<unsigned char> dyn(smem + 256);
    emu::g_dyn_smem = reinterpret_cast<unsigned char *>(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
    const unsigned tiles = (unsigned)((count + tile - 1) / tile);
    emu::launch(tiles, 256, [&] { dev::k_frontend(a); });
    size_t bad = 0, first_bad = 0;
    for (uint64_t i = 0; i < 2 * count; ++i) if (memcmp(&out[i], &want[i], 4) != 0) { if (!bad) first_bad = i; ++bad; }
    CHECK(bad == 0, "%s: %zu of %llu output floats differ from the oracle (first at %zu: %g vs %g)", c.name, bad,
          (unsigned long long)(2 * count), first_bad, out[first_bad], want[first_bad]);
    CHECK(out[2 * count] == -7.f, "%s: wrote past the end", c.name);
  }
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("emu_front seed %llu: equal (6 configurations)\n", (unsigned long long)seed);
  return 0;
}
