// k_spectrum.cu -- K8: cnr_fft (sdr.h:1273-1345) and spectrum (sdr.h:1347-1404).
//
// Both runnables look at one block (4096 / 1024 samples) of the stream in front of the FIR
// once per `decimation` samples (1 Hz, leandvb.cc:328,341), run cfft_engine::inplace(reverse)
// (dsp.h:78-110) on it, low-pass the power spectrum with a one-pole average over the
// measurements and reduce it to a C/N figure or a 1024-bin row.
//   k_meas_power  one CTA per measured block: load (format conversion, scaling and the
//                 rotator of leandvb.cc:310-318 applied on the fly -- the rotated stream is
//                 never materialised, the front-end kernel fuses it too), bit-reversal,
//                 the radix-2 stages in the reference's butterfly order with the host-built
//                 twiddles, 1/n, |.|^2.  Parallel over blocks and butterflies; every element
//                 sees the same operations as in the serial loop, so it is bit-identical.
//   k_meas_ema    the average is a recurrence over measurements: one CTA walks them in
//                 order; the band sums of avgslots() are serial float sums in the
//                 reference's index order (one thread), the divisions are IEEE.
// logf / log10f are evaluated on the host (glibc), like every other libm call of the path.
#include "common.cuh"
#include "kernels.h"

namespace ldvb {

namespace {

__device__ __forceinline__ float2 meas_load(const MeasSrc &s, uint64_t idx) {
  float2 v;
  if (idx < s.carry_count) {
    v = s.carry[idx];
  } else {
    uint64_t i = idx - s.carry_count + s.rest_off;
    const void *raw = s.rest.head;
    if (s.rest.main && i >= s.rest.c0) { raw = s.rest.main; i -= s.rest.c0; }
    switch (s.fmt) {
      case 0: { uchar2 q = reinterpret_cast<const uchar2 *>(raw)[i]; v = make_float2((float)((int)q.x - 128), (float)((int)q.y - 128)); break; }
      case 1: { char2 q = reinterpret_cast<const char2 *>(raw)[i]; v = make_float2((float)(int)q.x, (float)(int)q.y); break; }
      case 2: { ushort2 q = reinterpret_cast<const ushort2 *>(raw)[i]; v = make_float2((float)((int)q.x - 32768), (float)((int)q.y - 32768)); break; }
      case 3: { short2 q = reinterpret_cast<const short2 *>(raw)[i]; v = make_float2((float)(int)q.x, (float)(int)q.y); break; }
      case 4: { float2 q = __ldg(reinterpret_cast<const float2 *>(raw) + i); v = make_float2(fmul(q.x, s.scale), fmul(q.y, s.scale)); break; }
      default: v = __ldg(reinterpret_cast<const float2 *>(raw) + i); break;
    }
  }
  return v;
}

__device__ __forceinline__ float2 meas_rotate(const MeasSrc &s, uint64_t idx, float2 v) {
  if (!s.rot_lut) return v;
  const uint32_t ri = (uint32_t)(s.rot_index0 + idx) & 0xffffu;   // sdr.h:1246-1254
  const float c = __ldg(s.rot_lut + ri), sn = __ldg(s.rot_lut + 65536 + ri);
  return make_float2(fsub(fmul(v.x, c), fmul(v.y, sn)), fadd(fmul(v.x, sn), fmul(v.y, c)));
}

__global__ void __launch_bounds__(1024)
k_meas_power(MeasArgs a) {
  __shared__ float2 d[4096];
  const int n = 1 << a.logn, tid = threadIdx.x;
  const uint64_t base = a.point_start[blockIdx.x];
  for (int i = tid; i < n; i += 1024) {
    const int r = (int)(__brev((unsigned)i) >> (32 - a.logn));   // dsp.h:79-83
    d[r] = meas_rotate(a.src, base + i, meas_load(a.src, base + i));
  }
  __syncthreads();
  const int tw_stride = 4096 >> a.logn;   // omega_rev of an n-point engine = every (4096/n)-th entry
  for (int s = 0; s < a.logn; ++s) {      // dsp.h:85-102
    const int hbs = 1 << s, dom = 1 << (a.logn - 1 - s);
    for (int b = tid; b < n / 2; b += 1024) {
      const int j = b >> s, k = b & (hbs - 1);
      const int pidx = j * hbs * 2 + k, qidx = pidx + hbs;
      const float2 w = a.twiddle_rev[k * dom * tw_stride];
      const float2 q = d[qidx], p = d[pidx];
      const float xr = fsub(fmul(w.x, q.x), fmul(w.y, q.y));
      const float xi = fadd(fmul(w.x, q.y), fmul(w.y, q.x));
      d[qidx] = make_float2(fsub(p.x, xr), fsub(p.y, xi));
      d[pidx] = make_float2(fadd(p.x, xr), fadd(p.y, xi));
    }
    __syncthreads();
  }
  const float invn = 1.0f / (float)n;     // dsp.h:104-109 (exact: n is a power of two)
  float *out = a.power + (size_t)blockIdx.x * n;
  for (int i = tid; i < n; i += 1024) {
    const float re = fmul(d[i].x, invn), im = fmul(d[i].y, invn);
    out[i] = fadd(fmul(re, re), fmul(im, im));   // sdr.h:1313, 1380
  }
}

__device__ float meas_avgslots(const float *avg, int n, int i0, int i1) {   // sdr.h:1333-1337
  float s = 0;
  for (int i = i0; i <= i1; ++i) s = fadd(s, avg[i & (n - 1)]);
  return __fdiv_rn(s, (float)(i1 - i0 + 1));
}

__global__ void __launch_bounds__(1024)
k_meas_ema(MeasEmaArgs a) {
  const int tid = threadIdx.x, n = a.n;
  const float omk = fsub(1.0f, a.kavg);
  int have = *a.have;
  for (int p = 0; p < a.npoints; ++p) {
    const float *pw = a.power + (size_t)p * n;
    for (int i = tid; i < n; i += 1024) {
      float v = have ? a.avg[i] : pw[i];                       // "initialize with first spectrum"
      v = fadd(fmul(v, omk), fmul(pw[i], a.kavg));             // sdr.h:1321, 1388
      a.avg[i] = v;
      if (a.rows) a.rows[(size_t)p * n + i] = v;
    }
    have = 1;
    __syncthreads();
    if (a.bwslots && tid == 0) {                               // sdr.h:1323-1329
      const int bw = a.bwslots, icf = a.icf;
      a.sums[3 * p + 0] = meas_avgslots(a.avg, n, icf - bw, icf + bw);
      a.sums[3 * p + 1] = meas_avgslots(a.avg, n, icf - bw * 4, icf - bw * 3);
      a.sums[3 * p + 2] = meas_avgslots(a.avg, n, icf + bw * 3, icf + bw * 4);
    }
    __syncthreads();
  }
  if (tid == 0) *a.have = have;
}

__global__ void k_meas_save(MeasSrc src, uint64_t start, uint32_t count, float2 *dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = meas_load(src, start + i);
}

}  // namespace

cudaError_t launch_meas_power(const MeasArgs &a, cudaStream_t st) {
  if (a.npoints <= 0) return cudaSuccess;
  k_meas_power<<<a.npoints, 1024, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_meas_ema(const MeasEmaArgs &a, cudaStream_t st) {
  if (a.npoints <= 0) return cudaSuccess;
  k_meas_ema<<<1, 1024, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_meas_save(const MeasSrc &src, uint64_t start, uint32_t count, float2 *dst, cudaStream_t st) {
  if (!count) return cudaSuccess;
  k_meas_save<<<(count + 255) / 256, 256, 0, st>>>(src, start, count, dst);
  return cudaGetLastError();
}

}  // namespace ldvb
