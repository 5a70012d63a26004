// k_frontend.cu -- K1: format conversion / scaling -> frequency shift ->
// FIR low-pass with decimation, fused in one pass over the IQ stream.
//
// Replaces, per output sample and in this order, the reference runnables
//   cconverter<T,Z,f32,0,1,1>   dsp.h:33-54     f = (float)((int)v - Z)
//   scaler<float,cf32,cf32>     dsp.h:140-160   f = v * scale
//   rotator<f32>                sdr.h:1242-1254 (re*c - im*s, re*s + im*c), index & 0xffff
//   fir_filter<cf32,float>      dsp.h:246-259   y[k] = sum_i sc[i] * x[k*D + N - i]
//   decimator<cf32>             generic.h:254-261 (N == 0): y[k] = x[k*D]
// so the intermediate pipebufs (p_rawiq, p_derot) never touch HBM.
//
// Bit-exactness: the sum runs i = 0..N-1 from an accumulator of (0,0), each
// complex product as (c.re*x.re - c.im*x.im, c.re*x.im + c.im*x.re), no FMA.
//
// Data movement: one CTA = one tile of TILE_OUT outputs.  The raw input span of
// the tile (TILE_OUT*D + N samples, 2..8 bytes each) is brought into shared
// memory by ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) signalled on an
// mbarrier; conversion/rotation happen in shared memory; each thread then owns
// outputs t, t+256, ... so that both the shared-memory reads (consecutive
// float2) and the global stores (consecutive float2, streaming) are coalesced.
// Algorithmic HBM bytes per input sample: bytes_in + 8/D  (SURVEY.md 8d).
#include "common.cuh"
#include "kernels.h"

namespace ldvb {

namespace {

constexpr int kThreads = 256;
constexpr int kOutPerThread = 16;

__device__ __forceinline__ float2 convert_raw(const unsigned char *raw, int fmt, uint32_t idx,
                                              float scale) {
  switch (fmt) {
    case 0: {  // u8, zero at 128
      uchar2 v = reinterpret_cast<const uchar2 *>(raw)[idx];
      return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128));
    }
    case 1: {  // s8
      char2 v = reinterpret_cast<const char2 *>(raw)[idx];
      return make_float2((float)(int)v.x, (float)(int)v.y);
    }
    case 2: {  // u16, zero at 32768
      ushort2 v = reinterpret_cast<const ushort2 *>(raw)[idx];
      return make_float2((float)((int)v.x - 32768), (float)((int)v.y - 32768));
    }
    case 3: {  // s16
      short2 v = reinterpret_cast<const short2 *>(raw)[idx];
      return make_float2((float)(int)v.x, (float)(int)v.y);
    }
    case 4: {  // f32 through scaler
      float2 v = reinterpret_cast<const float2 *>(raw)[idx];
      return make_float2(fmul(v.x, scale), fmul(v.y, scale));
    }
    default: {  // 5: cf32 already preprocessed (after the notch), no scaling
      return reinterpret_cast<const float2 *>(raw)[idx];
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_frontend(FrontendArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;

  const uint32_t N = a.ntaps, D = a.decim;
  const uint32_t tile_out = a.tile_out;
  const uint64_t k0 = (uint64_t)blockIdx.x * tile_out;
  if (k0 >= a.count) return;
  const uint32_t nout = (uint32_t)min((uint64_t)tile_out, a.count - k0);
  // First / last input sample touched by this tile.
  const uint64_t g_first = k0 * D + (N ? 1 : 0);
  const uint32_t span = (nout - 1) * D + (N ? N : 1);
  // Which part of the two-part stream holds this tile (kernels.h: RawSrc).
  const unsigned char *part = static_cast<const unsigned char *>(a.src.head);
  uint64_t g_rel = g_first;
  if (a.src.main && g_first >= a.src.c0) {
    part = static_cast<const unsigned char *>(a.src.main);
    g_rel = g_first - a.src.c0;
  }
  // Align the bulk copy down to 16 bytes.
  const uint32_t bps = a.bytes_per_sample;
  const uint32_t align_elems = 16 / bps;
  const uint64_t g_al = g_rel & ~(uint64_t)(align_elems - 1);
  const uint32_t lead = (uint32_t)(g_rel - g_al);
  const uint32_t bytes = ((span + lead) * bps + 15u) & ~15u;

  // Shared memory map: [taps N*8][raw bytes][cf32 (only when raw is not cf32 in place)]
  float2 *s_taps = reinterpret_cast<float2 *>(smem);
  unsigned char *s_raw = smem + (((size_t)N * 8 + 127) & ~(size_t)127);
  const bool in_place = (a.fmt >= 4);
  float2 *s_x = in_place ? reinterpret_cast<float2 *>(s_raw)
                         : reinterpret_cast<float2 *>(s_raw + (((size_t)a.max_raw_bytes + 127) & ~(size_t)127));

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_1d(s_raw, part + g_al * bps, bytes, &bar);
  }
  for (uint32_t i = threadIdx.x; i < N; i += kThreads) s_taps[i] = a.taps[i];
  mbar_wait(&bar, 0);

  // Conversion / scaling / rotation in shared memory.
  const bool need_pass = (a.fmt != 5) || (a.rot_lut != nullptr);
  if (need_pass && !(a.fmt == 4 && a.scale == 1.0f && a.rot_lut == nullptr)) {
    const uint32_t total = span + lead;
    for (uint32_t i = threadIdx.x; i < total; i += kThreads) {
      float2 v = convert_raw(s_raw, a.fmt, i, a.scale);
      if (a.rot_lut) {
        uint32_t ri = (uint32_t)(a.rot_index0 + (g_first - lead) + i) & 0xffffu;
        float c = __ldg(a.rot_lut + ri), s = __ldg(a.rot_lut + 65536 + ri);
        v = make_float2(fsub(fmul(v.x, c), fmul(v.y, s)), fadd(fmul(v.x, s), fmul(v.y, c)));
      }
      if (!in_place) s_x[i] = v;
      else reinterpret_cast<float2 *>(s_raw)[i] = v;  // each thread rewrites its own element
    }
  }
  __syncthreads();

  const float2 *x = s_x + lead;  // x[j] = input sample g_first + j
  float2 *out = a.out + k0;
  if (N == 0) {
    for (uint32_t o = threadIdx.x; o < nout; o += kThreads) st_stream(out + o, x[(size_t)o * D]);
    return;
  }
  if (D == 1) {
    // Undecimated filter: each thread produces PAIRS of consecutive outputs so that the
    // N+1 input samples of a pair are read once (sliding window), and -- when every tap is
    // real, which is the case whenever the filter is not retuned (dsp.h:270-280 with f=0:
    // im = c*sinf(0) = +-0) -- the products with the zero imaginary part are skipped.
    // That is bit-exact for finite inputs: a term c.im*x = +-0 can only change the sign of
    // an exact zero, and an accumulator that starts at +0 never becomes -0.
    const uint32_t t2 = 2 * threadIdx.x;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
    for (uint32_t o = t2; o < nout; o += 2 * kThreads) {
      const float2 *xs = x + o;
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      float2 prev = xs[N];
      if (a.real_taps) {
#pragma unroll 5
        for (uint32_t i = 0; i < N; ++i) {
          const float2 cur = xs[N - 1 - i];
          const float c = s_taps[i].x;
          a1.x = fadd(a1.x, fmul(c, prev.x)); a1.y = fadd(a1.y, fmul(c, prev.y));
          a0.x = fadd(a0.x, fmul(c, cur.x));  a0.y = fadd(a0.y, fmul(c, cur.y));
          prev = cur;
        }
      } else {
        for (uint32_t i = 0; i < N; ++i) {
          const float2 cur = xs[N - 1 - i];
          const float2 c = s_taps[i];
          const float2 p1 = cmul(c, prev), p0 = cmul(c, cur);
          a1.x = fadd(a1.x, p1.x); a1.y = fadd(a1.y, p1.y);
          a0.x = fadd(a0.x, p0.x); a0.y = fadd(a0.y, p0.y);
          prev = cur;
        }
      }
      if (o + 1 < nout) {
        if (aligned16) __stcs(reinterpret_cast<float4 *>(out + o), make_float4(a0.x, a0.y, a1.x, a1.y));
        else { st_stream(out + o, a0); st_stream(out + o + 1, a1); }
      } else {
        st_stream(out + o, a0);
      }
    }
    return;
  }
  // Output o (local) = sum_i taps[i] * x[o*D + (N-1) - i].
  float2 acc[kOutPerThread];
#pragma unroll
  for (int r = 0; r < kOutPerThread; ++r) acc[r] = make_float2(0.f, 0.f);
  const uint32_t t = threadIdx.x;
  for (uint32_t i = 0; i < N; ++i) {
    const float2 c = s_taps[i];
    const uint32_t base = (N - 1) - i;
#pragma unroll
    for (int r = 0; r < kOutPerThread; ++r) {
      const uint32_t o = t + r * kThreads;
      if (o < nout) {
        const float2 p = cmul(c, x[(size_t)o * D + base]);
        acc[r].x = fadd(acc[r].x, p.x);
        acc[r].y = fadd(acc[r].y, p.y);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kOutPerThread; ++r) {
    const uint32_t o = t + r * kThreads;
    if (o < nout) st_stream(out + o, acc[r]);
  }
}

}  // namespace

int frontend_bytes_per_sample(int fmt) {
  switch (fmt) {
    case 0: case 1: return 2;
    case 2: case 3: return 4;
    default: return 8;
  }
}

cudaError_t launch_frontend(FrontendArgs a, cudaStream_t st) {
  if (a.count == 0) return cudaSuccess;
  const uint32_t N = a.ntaps, D = a.decim ? a.decim : 1;
  a.decim = D;
  a.bytes_per_sample = (uint32_t)frontend_bytes_per_sample(a.fmt);
  // Tile: up to 4096 outputs, bounded so that the staged input stays <= 8192 samples.
  uint32_t tile = kThreads * kOutPerThread;
  const uint32_t max_in = 8192;
  while (tile > kThreads && (uint64_t)(tile - 1) * D + N + 16 > max_in) tile -= kThreads;
  if ((uint64_t)(tile - 1) * D + N + 16 > 3 * max_in) return cudaErrorInvalidValue;  // taps too long
  a.tile_out = tile;
  const uint32_t span_max = (tile - 1) * D + (N ? N : 1) + 16;
  a.max_raw_bytes = (span_max * a.bytes_per_sample + 15u) & ~15u;
  size_t smem = (((size_t)N * 8 + 127) & ~(size_t)127) + (((size_t)a.max_raw_bytes + 127) & ~(size_t)127);
  if (a.fmt < 4) smem += (size_t)span_max * 8;
  static PerDeviceMark configured;   // per device: a process may hold handles on several GPUs
  if (smem > 48 * 1024 && configured.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(k_frontend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.commit(smem);
  }
  const uint64_t tiles = (a.count + tile - 1) / tile;
  k_frontend<<<(unsigned)tiles, kThreads, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace ldvb
