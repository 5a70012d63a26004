// tx.cu -- B200-native leandvbtx transmit chain: kernels, handle and C ABI
// (include/leandvb_b200_tx.h).
//
// Replaces the runnables of apps/leandvbtx.cc:79-197 (citations: file:line under
// /root/reference/src/leansdr/):
//   randomizer dvb.h:1063-1102, rs_encoder dvb.h:957-980 + rs.h:141-167,
//   interleaver dvb.h:900-921, dvb_convol dvb.h:519-604 + convolutional.h:225-270,
//   cstln_transmitter sdr.h:1196-1221, fir_resampler dsp.h:290-364,
//   decimator generic.h:247-267, simple_agc sdr.h:238-274.
//
// Every stage is position-deterministic in the reference (no feedback from the
// receiver side), so all of them are data parallel except the AGC's one-pole
// estimate, which advances once per 128-sample chunk and is walked serially by one
// lane over per-chunk powers that were summed in parallel (in the reference's
// order).  Streams between stages are flat device buffers "[carry | new]" like on
// the receive side; all counts are known on the host before anything is launched,
// so a call makes no device->host read.
//
//   k_tx_rs         warp per packet: XOR with the PRBS pattern, RS(204,188) parity by a
//                   16-lane LFSR (lane i holds the register of X^(15-i))
//   k_tx_interleave thread per byte: row r, byte i <- packet r + 11 - i%12
//   k_tx_convol     thread per puncturing group: 16-bit history gathered from 3 bytes,
//                   bits_out parities, bits_out/bps symbols
//   k_tx_resample   thread per OUTPUT sample (after the decimator): <= ceil(N/I) taps
//                   of the polyphase branch t%I, constellation points from a table
//                   (cstln_transmitter fused), interpolated samples the decimator
//                   would drop are never computed
//   k_tx_amp2 / k_tx_agc / k_tx_scale   simple_agc
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/leandvb_b200.h"
#include "../../include/leandvb_b200_tx.h"
#include "common.cuh"
#include "tables.h"

namespace {

using namespace ldvb;

constexpr int kAgcChunk = 128;     // simple_agc::chunk_size (sdr.h:252)
constexpr int kInterDepth = 12;    // interleaver branches (dvb.h:908)

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t alloc(size_t n) { bytes = n; return cudaMalloc(&p, n ? n : 16); }
  void release() { if (p) cudaFree(p); p = nullptr; }
  template <class T> T *as() const { return static_cast<T *>(p); }
};

// ------------------------------------------------------------------ kernels

// leantsgen (apps/leantsgen.cc:37-47): byte 4k = 4k, bytes 4k+1..4k+3 = 24-bit packet counter,
// byte 0 = 0x47.
__global__ void k_tx_tsgen(uint64_t first, uint64_t n, uint8_t *ts) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 32-bit word each
  if (i >= n * 47) return;
  const uint64_t p = i / 47; const uint32_t w = (uint32_t)(i % 47);
  const uint32_t t = (uint32_t)(first + p);
  uint32_t b0 = w ? 4u * w : 0x47u;
  const uint32_t word = b0 | (((t >> 16) & 0xffu) << 8) | (((t >> 8) & 0xffu) << 16) | ((t & 0xffu) << 24);
  reinterpret_cast<uint32_t *>(ts)[i] = word;   // 188 = 47 words: packets stay 4-byte aligned
}

struct TxRsArgs {
  const uint8_t *ts;        // [n][188]
  uint8_t *rs;              // [n][204] (already offset past the carried packets)
  uint64_t first_packet;    // absolute index of ts[0]: PRBS position = (index % 8) * 188
  uint32_t n;
  const uint8_t *pattern;   // 1504 bytes (dvb.h:1073-1085)
  const uint8_t *gf_exp;    // 512
  const uint8_t *gf_log;    // 256
  const uint8_t *g;         // G[0..16], G[0] = leading coefficient (rs.h:91-102)
};

__global__ void __launch_bounds__(128) k_tx_rs(TxRsArgs a) {
  __shared__ uint8_t s_exp[512];
  __shared__ uint8_t s_log[256];
  __shared__ uint8_t s_pkt[4][192];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s_exp[i] = a.gf_exp[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_log[i] = a.gf_log[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t p = blockIdx.x * 4 + warp;
  const bool have = p < a.n;
  if (have) {
    const uint8_t *src = a.ts + (size_t)p * 188;
    const uint8_t *pat = a.pattern + ((a.first_packet + p) & 7u) * 188;
    uint8_t *dst = a.rs + (size_t)p * 204;
    for (int i = lane; i < 188; i += 32) {
      const uint8_t v = src[i] ^ pat[i];               // randomizer (dvb.h:1093)
      s_pkt[warp][i] = v;
      dst[i] = v;                                      // message part (dvb.h:969)
    }
  }
  __syncthreads();
  if (!have) return;
  // Remainder of P(X)*X^16 modulo G (rs.h:152-160): lane i < 16 holds the pending
  // correction of coefficient d+1+i; k = p[d] / G[0], register i <- register i+1 ^ k*G[i+1].
  const int gi = (lane < 16) ? a.g[lane + 1] : 0;
  const int lg = gi ? s_log[gi] : -1;
  const int lg0 = s_log[a.g[0]];
  uint32_t reg = 0;
  for (int d = 0; d < 188; ++d) {
    const uint32_t r0 = __shfl_sync(0xffffffffu, reg, 0);
    const uint32_t cur = s_pkt[warp][d] ^ r0;
    const uint32_t up = __shfl_down_sync(0xffffffffu, reg, 1);
    uint32_t k = 0;
    if (cur) k = s_exp[s_log[cur] + 255 - lg0];        // gf.div (rs.h:68-72)
    uint32_t m = 0;
    if (k && lg >= 0) m = s_exp[s_log[k] + lg];        // gf.mul (rs.h:64-67)
    reg = ((lane < 15) ? up : 0u) ^ m;
  }
  if (lane < 16) a.rs[(size_t)p * 204 + 188 + lane] = (uint8_t)reg;
}

// interleaver::run (dvb.h:908-912): out row r byte i = packet (r + 11 - i % 12) byte i.
__global__ void k_tx_interleave(const uint8_t *rs, uint64_t rows, uint8_t *out) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 204) return;
  const uint64_t r = idx / 204; const uint32_t i = (uint32_t)(idx % 204);
  out[idx] = rs[(r + (kInterDepth - 1) - i % kInterDepth) * 204 + i];
}

struct TxConvArgs {
  const uint8_t *bytes;     // [2 history bytes | bytes to encode]
  uint64_t ngroups;
  int bits_in, bits_out, bps;
  uint16_t polys[8];
  uint8_t *sym;             // ngroups * bits_out / bps symbols
};

// convol_multipoly::encode (convolutional.h:236-262): after input bit k the history holds
// bits k-15..k with bit k in position 15; a group of bits_in input bits yields bits_out
// parities (MSB first), cut into bps-bit symbols.
__global__ void k_tx_convol(TxConvArgs a) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.ngroups) return;
  const uint64_t ke = (g + 1) * (uint64_t)a.bits_in - 1;     // last input bit of the group
  const uint64_t byte = (ke >> 3) + 2;                       // + the two history bytes
  const uint32_t w = ((uint32_t)a.bytes[byte - 2] << 16) | ((uint32_t)a.bytes[byte - 1] << 8) | a.bytes[byte];
  const uint32_t v = (w >> (7 - (uint32_t)(ke & 7))) & 0xffffu;   // LSB = newest bit
  const uint32_t hist = __brev(v) >> 16;                          // newest bit in position 15
  uint32_t ser = 0;
  for (int p = 0; p < a.bits_out; ++p) ser = (ser << 1) | (__popc(hist & a.polys[p]) & 1u);
  const int nsym = a.bits_out / a.bps;
  const uint32_t mask = (1u << a.bps) - 1u;
  uint8_t *out = a.sym + g * (uint64_t)nsym;
  for (int s = 0; s < nsym; ++s) out[s] = (uint8_t)((ser >> (a.bits_out - (s + 1) * a.bps)) & mask);
}

struct TxResampleArgs {
  const uint8_t *sym;       // symbols from absolute index n0 on (SYMBOLS source)
  const float2 *x;          // or cf32 input from absolute index n0 on (stand-alone resampler)
  const float2 *points;     // [nsymbols] constellation points as floats (sdr.h:1212-1214)
  const float2 *taps;       // [ncoeffs] shifted_coeffs (dsp.h:352-361)
  int ncoeffs, interp, decim, latency;
  uint64_t n0;              // absolute index of sym[0] / x[0]
  uint64_t m0;              // absolute index (after the decimator) of out[0]
  uint64_t count;           // outputs
  float2 *out;
};

// fir_resampler::run (dsp.h:325-336) followed by decimator::run (generic.h:256-261):
// decimated sample m is interpolated sample t = m*decim = n*interp + p,
//   y = sum_j taps[p + j*interp] * x[n + latency - j], accumulated from 0 in that order.
template <bool FROM_SYMBOLS>
__global__ void __launch_bounds__(256) k_tx_resample(TxResampleArgs a) {
  extern __shared__ float2 s_taps[];
  for (int i = threadIdx.x; i < a.ncoeffs; i += blockDim.x) s_taps[i] = a.taps[i];
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  const uint64_t t = (a.m0 + i) * (uint64_t)a.decim;
  const uint64_t n = t / (uint64_t)a.interp;
  const int p = (int)(t - n * (uint64_t)a.interp);
  int64_t xi = (int64_t)(n - a.n0) + a.latency;
  float xr = 0.f, xim = 0.f;
  for (int c = p; c < a.ncoeffs; c += a.interp, --xi) {
    const float2 tc = s_taps[c];
    float2 v;
    if (FROM_SYMBOLS) v = a.points[a.sym[xi]]; else v = a.x[xi];
    const float2 pr = cmul(tc, v);                       // (*pc)*(*pi), math.h:38-41
    xr = fadd(xr, pr.x); xim = fadd(xim, pr.y);
  }
  a.out[i] = make_float2(xr, xim);
}

// simple_agc::run (sdr.h:256-259): amp2 = (sum over the chunk, in order, of re*re + im*im) / 128.
__global__ void k_tx_amp2(const float2 *x, uint64_t nchunks, float *amp2) {
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const float4 *p = reinterpret_cast<const float4 *>(x + c * kAgcChunk);
  float acc = 0.f;
#pragma unroll 8
  for (int i = 0; i < kAgcChunk / 2; ++i) {
    const float4 v = p[i];
    acc = fadd(acc, fadd(fmul(v.x, v.x), fmul(v.y, v.y)));
    acc = fadd(acc, fadd(fmul(v.z, v.z), fmul(v.w, v.w)));
  }
  amp2[c] = __fdiv_rn(acc, (float)kAgcChunk);
}

// simple_agc::run (sdr.h:260-262): the estimate advances once per chunk.  One CTA: the threads
// stage 1024 chunk powers (and their products with bw) in shared memory, thread 0 walks the
// two-operation chain est = est*(1-bw) + amp2*bw over them, the threads turn the 1024 estimates
// into gains.  `if (!estimated) estimated = amp2` can only fire while the estimate is zero; once
// it is a normal positive number it stays one (amp2 >= 0, 1-bw >= 1/2), so the test leaves the chain.
constexpr int kAgcTile = 1024;
__global__ void __launch_bounds__(256) k_tx_agc(const float *amp2, uint64_t nchunks, float bw, float out_rms,
                                                float *estimated_io, float *gain) {
  __shared__ float s_a[kAgcTile], s_ab[kAgcTile], s_e[kAgcTile];
  __shared__ float s_est;
  const float omb = fsub(1.0f, bw);
  const bool safe = bw <= 0.5f && bw >= 0.0f;
  if (threadIdx.x == 0) s_est = *estimated_io;
  for (uint64_t base = 0; base < nchunks; base += kAgcTile) {
    const int m = (int)min((uint64_t)kAgcTile, nchunks - base);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const float a = amp2[base + i];
      s_a[i] = a; s_ab[i] = fmul(a, bw);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float est = s_est;
      int i = 0;
      while (i < m) {
        if (safe && est >= 1e-30f && est <= 3e38f && i + 16 <= m) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = s_ab[i + j];           // loads off the chain
#pragma unroll
          for (int j = 0; j < 16; ++j) { est = fadd(fmul(est, omb), v[j]); v[j] = est; }
#pragma unroll
          for (int j = 0; j < 16; ++j) s_e[i + j] = v[j];
          i += 16;
        } else {
          if (est == 0.0f) est = s_a[i];
          est = fadd(fmul(est, omb), s_ab[i]);
          s_e[i] = est; ++i;
        }
      }
      s_est = est;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const float e = s_e[i];
      gain[base + i] = (e != 0.0f) ? __fdiv_rn(out_rms, __fsqrt_rn(e)) : 0.f;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *estimated_io = s_est;
}

__global__ void k_tx_scale(const float2 *x, const float *gain, uint64_t n, float2 *out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = gain[i / kAgcChunk];
  const float2 v = x[i];
  out[i] = make_float2(fmul(v.x, g), fmul(v.y, g));      // sdr.h:266-269
}

// ------------------------------------------------------------------- tables

// dvb.h:520-565: fec_specs.
bool tx_fec_spec(int fec, int *bits_in, int *bits_out, uint16_t polys[8]) {
  const uint16_t G1 = 0171, G2 = 0133;                   // dvb.h:84-85
  memset(polys, 0, 16);
  switch (fec) {
    case LDVB_FEC12: polys[0] = G1; polys[1] = G2; *bits_in = 1; *bits_out = 2; return true;
    case LDVB_FEC23: polys[0] = G1; polys[1] = G2; polys[2] = G2 << 1; *bits_in = 2; *bits_out = 3; return true;
    case LDVB_FEC46: polys[0] = G1; polys[1] = G2; polys[2] = G2 << 1; polys[3] = G1 << 2; polys[4] = G2 << 2;
      polys[5] = G2 << 3; *bits_in = 4; *bits_out = 6; return true;
    case LDVB_FEC34: polys[0] = G1; polys[1] = G2; polys[2] = G2 << 1; polys[3] = G1 << 2; *bits_in = 3; *bits_out = 4;
      return true;
    case LDVB_FEC56: polys[0] = G1; polys[1] = G2; polys[2] = G2 << 1; polys[3] = G1 << 2; polys[4] = G2 << 3;
      polys[5] = G1 << 4; *bits_in = 5; *bits_out = 6; return true;
    case LDVB_FEC78: polys[0] = G1; polys[1] = G2; polys[2] = G2 << 1; polys[3] = G2 << 2; polys[4] = G2 << 3;
      polys[5] = G1 << 4; polys[6] = G2 << 5; polys[7] = G1 << 6; *bits_in = 7; *bits_out = 8; return true;
    default: return false;
  }
}

// rs_engine::rs_engine (rs.h:91-102): G = prod (X - alpha^d), G[0] = leading coefficient.
void tx_rs_generator(const uint8_t ex[512], const uint8_t lg[256], uint8_t G[17]) {
  auto mul = [&](uint8_t x, uint8_t y) -> uint8_t { return (x && y) ? ex[lg[x] + lg[y]] : 0; };
  for (int i = 0; i <= 16; ++i) G[i] = (i == 16) ? 1 : 0;
  for (int d = 0; d < 16; ++d)
    for (int i = 0; i <= 16; ++i) G[i] = (uint8_t)(((i == 16) ? 0 : G[i + 1]) ^ mul(ex[d], G[i]));
}

float tx_amp(const char *power_db) {                     // leandvbtx.cc:289
  char buf[33]; memcpy(buf, power_db, 32); buf[32] = 0;
  return expf(logf(10) * atof(buf) / 20);
}

}  // namespace

struct ldvbtx_handle {
  ldvbtx_config cfg;
  cudaStream_t st = nullptr;
  bool own_stream = true;
  std::string err;
  int fec = 0, bps = 0, bits_in = 0, bits_out = 0, nsymbols = 0;
  uint16_t polys[8];
  std::vector<float> taps;          // real taps (leandvbtx.cc:131-138)
  int ncoeffs = 0, latency = 0;
  float amp = 1, out_rms = 1, bw = 0.001f;
  uint64_t max_sym = 0, max_out = 0;
  DevBuf d_pattern, d_gfexp, d_gflog, d_g, d_points, d_sc;
  DevBuf d_rs, d_mb, d_sym, d_raw, d_amp2, d_gain, d_est, d_tmp, d_ts_in, d_out;
  // stream positions (see the file comment)
  uint64_t pk_total = 0;            // TS packets consumed so far
  uint32_t rs_carry = 0;            // packets kept at the front of d_rs (<= 11)
  uint64_t mb_left = 0;             // bytes kept behind the two history bytes of d_mb (< bits_in)
  uint64_t sym_count = 0;           // unread symbols at the front of d_sym
  uint64_t sym_n0 = 0;              // absolute index of d_sym[0] = input steps the resampler has done
  uint64_t dec_done = 0;            // samples the decimator has produced so far
  uint64_t raw_left = 0;            // samples kept at the front of d_raw (< 128)
  // taps of the last call
  uint64_t tap_rs_packets = 0, tap_mb_bytes = 0, tap_symbols = 0;
  DevBuf t_rs, t_mb, t_sym;
};

namespace {

#define TCK(call)                                                                     \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      char b_[256];                                                                   \
      snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      h->err = b_;                                                                    \
      return LDVB_ECUDA;                                                              \
    }                                                                                 \
  } while (0)

int tx_fail(ldvbtx_handle *h, int code, const char *msg) { h->err = msg; return code; }

// What one call produces, from the counters alone (no device access).
struct TxPlan {
  uint64_t rows, mb_avail, mb_used, groups, new_sym, sym_avail, steps, new_dec, raw_avail, chunks, n_out;
};

TxPlan tx_plan(const ldvbtx_handle *h, uint64_t n_packets) {
  TxPlan p;
  const uint64_t held = h->rs_carry + n_packets;
  p.rows = held >= kInterDepth ? held - (kInterDepth - 1) : 0;           // dvb.h:906
  p.mb_avail = h->mb_left + p.rows * 204;
  p.mb_used = p.mb_avail / h->bits_in * h->bits_in;                      // dvb.h:591-593
  p.groups = p.mb_used * 8 / h->bits_in;
  p.new_sym = p.groups * (uint64_t)(h->bits_out / h->bps);
  p.sym_avail = h->sym_count + p.new_sym;
  const uint64_t I = (uint64_t)h->cfg.interp, D = (uint64_t)h->cfg.decim;
  p.steps = 0;
  if (p.sym_avail >= (uint64_t)h->ncoeffs && p.sym_avail * I >= (uint64_t)h->ncoeffs)   // dsp.h:307, 320
    p.steps = (p.sym_avail * I - h->ncoeffs) / I;                        // dsp.h:321-322
  const uint64_t interp_total = (h->sym_n0 + p.steps) * I;
  p.new_dec = interp_total / D - h->dec_done;                            // generic.h:256
  p.raw_avail = h->raw_left + p.new_dec;
  p.chunks = p.raw_avail / kAgcChunk;
  p.n_out = h->cfg.agc ? p.chunks * kAgcChunk : p.new_dec;
  return p;
}

void tx_reset(ldvbtx_handle *h) {
  h->pk_total = 0; h->rs_carry = 0; h->mb_left = 0; h->sym_count = 0; h->sym_n0 = 0;
  h->dec_done = 0; h->raw_left = 0;
  cudaMemsetAsync(h->d_mb.p, 0, 2, h->st);               // hist = 0 (convolutional.h:232)
  cudaMemsetAsync(h->d_est.p, 0, 4, h->st);              // estimated = 0 (sdr.h:246)
}

// Moves `bytes` from src to dst inside one buffer (regions may overlap) through d_tmp.
int tx_move(ldvbtx_handle *h, void *dst, const void *src, size_t bytes) {
  if (!bytes || dst == src) return LDVB_OK;
  if (bytes > h->d_tmp.bytes) return tx_fail(h, LDVB_ESTATE, "carry larger than the scratch buffer");
  TCK(cudaMemcpyAsync(h->d_tmp.p, src, bytes, cudaMemcpyDeviceToDevice, h->st));
  TCK(cudaMemcpyAsync(dst, h->d_tmp.p, bytes, cudaMemcpyDeviceToDevice, h->st));
  return LDVB_OK;
}

int tx_process(ldvbtx_handle *h, const uint8_t *ts_dev, uint64_t n, float2 *out, uint64_t cap, uint64_t *n_out) {
  const ldvbtx_config &c = h->cfg;
  if (n > c.max_packets) return tx_fail(h, LDVB_EOVERFLOW, "n_packets exceeds max_packets");
  const TxPlan p = tx_plan(h, n);
  if (p.n_out > cap) return tx_fail(h, LDVB_EOVERFLOW, "output buffer too small (see ldvbtx_max_samples)");
  int rc;
  // ---- randomizer + rs_encoder
  if (n) {
    TxRsArgs a;
    a.ts = ts_dev; a.rs = h->d_rs.as<uint8_t>() + (size_t)h->rs_carry * 204; a.first_packet = h->pk_total;
    a.n = (uint32_t)n; a.pattern = h->d_pattern.as<uint8_t>(); a.gf_exp = h->d_gfexp.as<uint8_t>();
    a.gf_log = h->d_gflog.as<uint8_t>(); a.g = h->d_g.as<uint8_t>();
    k_tx_rs<<<(unsigned)((n + 3) / 4), 128, 0, h->st>>>(a);
    TCK(cudaGetLastError());
    if (c.keep_taps) {
      TCK(cudaMemcpyAsync(h->t_rs.p, a.rs, n * 204, cudaMemcpyDeviceToDevice, h->st));
      h->tap_rs_packets = n;
    }
  }
  // ---- interleaver
  uint8_t *mb = h->d_mb.as<uint8_t>();
  if (p.rows) {
    k_tx_interleave<<<(unsigned)((p.rows * 204 + 255) / 256), 256, 0, h->st>>>(h->d_rs.as<uint8_t>(), p.rows,
                                                                              mb + 2 + h->mb_left);
    TCK(cudaGetLastError());
    if (c.keep_taps) TCK(cudaMemcpyAsync(h->t_mb.p, mb + 2 + h->mb_left, p.rows * 204, cudaMemcpyDeviceToDevice, h->st));
  }
  h->tap_mb_bytes = p.rows * 204;
  {
    const uint64_t held = h->rs_carry + n;
    const uint64_t keep = std::min<uint64_t>(held, kInterDepth - 1);
    if ((rc = tx_move(h, h->d_rs.p, h->d_rs.as<uint8_t>() + (held - keep) * 204, keep * 204))) return rc;
    h->rs_carry = (uint32_t)keep;
    h->pk_total += n;
  }
  // ---- dvb_convol
  uint8_t *sym = h->d_sym.as<uint8_t>();
  if (p.groups) {
    TxConvArgs a;
    a.bytes = mb; a.ngroups = p.groups; a.bits_in = h->bits_in; a.bits_out = h->bits_out; a.bps = h->bps;
    memcpy(a.polys, h->polys, sizeof a.polys);
    a.sym = sym + h->sym_count;
    k_tx_convol<<<(unsigned)((p.groups + 255) / 256), 256, 0, h->st>>>(a);
    TCK(cudaGetLastError());
    if (c.keep_taps) TCK(cudaMemcpyAsync(h->t_sym.p, a.sym, p.new_sym, cudaMemcpyDeviceToDevice, h->st));
  }
  h->tap_symbols = p.new_sym;
  if (p.mb_used) {
    // new history = the last two bytes consumed; unread bytes follow them
    const uint64_t left = p.mb_avail - p.mb_used;
    if ((rc = tx_move(h, mb, mb + p.mb_used, 2 + left))) return rc;
    h->mb_left = left;
  } else {
    h->mb_left = p.mb_avail;
  }
  // ---- cstln_transmitter + fir_resampler + decimator
  float2 *raw = h->d_raw.as<float2>();
  float2 *res_dst = c.agc ? raw + h->raw_left : out;
  if (p.new_dec) {
    TxResampleArgs a;
    a.sym = sym; a.x = nullptr; a.points = h->d_points.as<float2>(); a.taps = h->d_sc.as<float2>();
    a.ncoeffs = h->ncoeffs; a.interp = c.interp; a.decim = c.decim; a.latency = h->latency;
    a.n0 = h->sym_n0; a.m0 = h->dec_done; a.count = p.new_dec; a.out = res_dst;
    k_tx_resample<true><<<(unsigned)((p.new_dec + 255) / 256), 256, (size_t)h->ncoeffs * 8, h->st>>>(a);
    TCK(cudaGetLastError());
  }
  if (p.steps) {
    const uint64_t left = p.sym_avail - p.steps;
    if ((rc = tx_move(h, sym, sym + p.steps, left))) return rc;
    h->sym_count = left;
    h->sym_n0 += p.steps;
  } else {
    h->sym_count = p.sym_avail;
  }
  h->dec_done += p.new_dec;
  // ---- simple_agc
  if (c.agc) {
    if (p.chunks) {
      k_tx_amp2<<<(unsigned)((p.chunks + 127) / 128), 128, 0, h->st>>>(raw, p.chunks, h->d_amp2.as<float>());
      TCK(cudaGetLastError());
      k_tx_agc<<<1, 256, 0, h->st>>>(h->d_amp2.as<float>(), p.chunks, h->bw, h->out_rms, h->d_est.as<float>(),
                                    h->d_gain.as<float>());
      TCK(cudaGetLastError());
      const uint64_t ns = p.chunks * kAgcChunk;
      k_tx_scale<<<(unsigned)((ns + 255) / 256), 256, 0, h->st>>>(raw, h->d_gain.as<float>(), ns, out);
      TCK(cudaGetLastError());
    }
    const uint64_t left = p.raw_avail - p.chunks * kAgcChunk;
    if (p.chunks && (rc = tx_move(h, raw, raw + p.chunks * kAgcChunk, left * 8))) return rc;
    h->raw_left = left;
  }
  *n_out = p.n_out;
  return LDVB_OK;
}

}  // namespace

extern "C" {

void ldvbtx_config_default(ldvbtx_config *c) {           // leandvbtx.cc:69-76
  if (!c) return;
  memset(c, 0, sizeof *c);
  c->abi_version = LDVB_ABI_VERSION;
  c->constellation = LDVB_CSTLN_QPSK;
  c->fec = LDVB_FEC12;
  c->interp = 2; c->decim = 1;
  c->rolloff = 0.35f; c->rrc_rej = 10;
  strcpy(c->power_db, "0");
  c->agc = 0;
  c->device = 0;
  c->max_packets = 4096;
}

const char *ldvbtx_last_error(const ldvbtx_handle *h) { return h ? h->err.c_str() : ""; }

int ldvbtx_host_taps(const ldvbtx_config *cfg, float *dst, size_t cap, size_t *n) {
  if (!cfg || !n || cfg->interp < 1) return LDVB_EINVAL;
  const std::vector<float> t = design_tx_rrc(cfg->interp, cfg->rolloff, cfg->rrc_rej, tx_amp(cfg->power_db));
  *n = t.size();
  if (dst) {
    if (cap < t.size()) return LDVB_EOVERFLOW;
    memcpy(dst, t.data(), t.size() * 4);
  }
  return LDVB_OK;
}

int ldvbtx_destroy(ldvbtx_handle *h) {
  if (!h) return LDVB_OK;
  cudaSetDevice(h->cfg.device);
  if (h->st) cudaStreamSynchronize(h->st);
  DevBuf *bufs[] = {&h->d_pattern, &h->d_gfexp, &h->d_gflog, &h->d_g, &h->d_points, &h->d_sc, &h->d_rs, &h->d_mb,
                    &h->d_sym, &h->d_raw, &h->d_amp2, &h->d_gain, &h->d_est, &h->d_tmp, &h->d_ts_in, &h->d_out,
                    &h->t_rs, &h->t_mb, &h->t_sym};
  for (DevBuf *b : bufs) b->release();
  if (h->st && h->own_stream) cudaStreamDestroy(h->st);
  delete h;
  return LDVB_OK;
}

int ldvbtx_create(const ldvbtx_config *cfg, ldvbtx_handle **out) {
  if (!cfg || !out) return LDVB_EINVAL;
  *out = nullptr;
  if (cfg->abi_version != LDVB_ABI_VERSION) return LDVB_EINVAL;
  if (cfg->interp < 1 || cfg->decim < 1 || cfg->max_packets < 1 || cfg->rrc_rej <= 0) return LDVB_EINVAL;
  if (cfg->constellation < LDVB_CSTLN_BPSK || cfg->constellation > LDVB_CSTLN_256QAM) return LDVB_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) return LDVB_ENODEV;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return LDVB_ENODEV;
  if (prop.major < 10) return LDVB_ENODEV;               // kernels are built for sm_100a only
  ldvbtx_handle *h = new (std::nothrow) ldvbtx_handle();
  if (!h) return LDVB_ENOMEM;
  h->cfg = *cfg;
  h->cfg.power_db[sizeof h->cfg.power_db - 1] = 0;
  const Cstln cs = make_cstln(cfg->constellation, cfg->fec, false);
  h->nsymbols = cs.nsymbols;
  h->bps = 0; while ((1 << h->bps) < cs.nsymbols) ++h->bps;
  h->fec = cfg->fec;
  if (h->fec == LDVB_FEC23 && (cs.nsymbols == 4 || cs.nsymbols == 64)) h->fec = LDVB_FEC46;     // leandvbtx.cc:115-119
  if (cs.nsymbols == 0 || !tx_fec_spec(h->fec, &h->bits_in, &h->bits_out, h->polys) || h->bits_out % h->bps) {   // dvb.h:582-584
    delete h;
    return LDVB_EINVAL;
  }
  h->amp = tx_amp(h->cfg.power_db);
  h->taps = design_tx_rrc(cfg->interp, cfg->rolloff, cfg->rrc_rej, h->amp);
  h->ncoeffs = (int)h->taps.size();
  h->latency = (h->ncoeffs + cfg->interp) / cfg->interp;                 // dsp.h:323
  h->out_rms = h->amp / sqrtf((float)cfg->interp / cfg->decim);          // leandvbtx.cc:163
  h->bw = 0.001 * cfg->decim / cfg->interp;                              // leandvbtx.cc:165
  if ((size_t)h->ncoeffs * 8 > 48 * 1024) { delete h; return LDVB_EINVAL; }

  if (cudaSetDevice(cfg->device) != cudaSuccess) { delete h; return LDVB_ENODEV; }
  if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) { delete h; return LDVB_ECUDA; }
  const uint64_t P = cfg->max_packets;
  const uint64_t max_mb = (P + kInterDepth) * 204 + 16;
  h->max_sym = (max_mb * 8 / h->bits_in + 1) * (uint64_t)(h->bits_out / h->bps) + h->ncoeffs + 64;
  h->max_out = h->max_sym * cfg->interp / cfg->decim + 2 * kAgcChunk;
  const std::vector<uint8_t> pattern = make_derand_pattern();
  uint8_t ex[512], lg[256], G[17];
  make_rs_tables(ex, lg);
  tx_rs_generator(ex, lg, G);
  std::vector<float> points(2 * 256, 0.f);
  for (int s = 0; s < cs.nsymbols; ++s) { points[2 * s] = 0 + cs.sym_re[s]; points[2 * s + 1] = 0 + cs.sym_im[s]; }
  const std::vector<float> sc = shift_taps_resampler(h->taps, 0.f);      // set_freq(0), dsp.h:302
  const size_t tmp_bytes = std::max<size_t>({(size_t)kInterDepth * 204, (size_t)h->ncoeffs + 4096, (size_t)kAgcChunk * 8, 64});
  struct { DevBuf *b; size_t n; const void *init; } allocs[] = {
      {&h->d_pattern, pattern.size(), pattern.data()}, {&h->d_gfexp, 512, ex}, {&h->d_gflog, 256, lg}, {&h->d_g, 17, G},
      {&h->d_points, points.size() * 4, points.data()}, {&h->d_sc, sc.size() * 4, sc.data()},
      {&h->d_rs, (P + kInterDepth) * 204, nullptr}, {&h->d_mb, max_mb + 16, nullptr}, {&h->d_sym, h->max_sym, nullptr},
      {&h->d_raw, cfg->agc ? h->max_out * 8 : 16, nullptr}, {&h->d_amp2, (h->max_out / kAgcChunk + 1) * 4, nullptr},
      {&h->d_gain, (h->max_out / kAgcChunk + 1) * 4, nullptr}, {&h->d_est, 4, nullptr}, {&h->d_tmp, tmp_bytes, nullptr},
      {&h->t_rs, cfg->keep_taps ? P * 204 : 16, nullptr}, {&h->t_mb, cfg->keep_taps ? max_mb : 16, nullptr},
      {&h->t_sym, cfg->keep_taps ? h->max_sym : 16, nullptr}};
  for (auto &al : allocs) {
    if (al.b->alloc(al.n) != cudaSuccess) { ldvbtx_destroy(h); return LDVB_ENOMEM; }
    if (al.init && cudaMemcpyAsync(al.b->p, al.init, al.n, cudaMemcpyHostToDevice, h->st) != cudaSuccess) {
      ldvbtx_destroy(h);
      return LDVB_ECUDA;
    }
  }
  tx_reset(h);
  if (cudaStreamSynchronize(h->st) != cudaSuccess) { ldvbtx_destroy(h); return LDVB_ECUDA; }   // host tables are locals
  *out = h;
  return LDVB_OK;
}

int ldvbtx_reset(ldvbtx_handle *h) {
  if (!h) return LDVB_EINVAL;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return tx_fail(h, LDVB_ECUDA, "cudaSetDevice");
  tx_reset(h);
  return LDVB_OK;
}

int ldvbtx_set_stream(ldvbtx_handle *h, void *cuda_stream) {
  if (!h) return LDVB_EINVAL;
  TCK(cudaStreamSynchronize(h->st));
  if (h->own_stream && h->st) cudaStreamDestroy(h->st);
  h->st = static_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  return LDVB_OK;
}

size_t ldvbtx_max_samples(const ldvbtx_handle *h, size_t n_packets) {
  if (!h) return 0;
  // the plan from the current state, plus what an unknown state could add: one interleaver
  // fill, the filter history and a partial AGC chunk
  const uint64_t sym = ((n_packets + kInterDepth) * 204 * 8 / h->bits_in + 1) * (uint64_t)(h->bits_out / h->bps) + h->ncoeffs;
  return (size_t)(sym * h->cfg.interp / h->cfg.decim + 2 * kAgcChunk);
}

int ldvbtx_process_device(ldvbtx_handle *h, const uint8_t *ts_dev, size_t n_packets, float *iq_dev, size_t cap_samples,
                          size_t *n_samples) {
  if (!h || !n_samples || (!ts_dev && n_packets) || !iq_dev) return LDVB_EINVAL;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return tx_fail(h, LDVB_ECUDA, "cudaSetDevice");
  uint64_t got = 0;
  const int rc = tx_process(h, ts_dev, n_packets, reinterpret_cast<float2 *>(iq_dev), cap_samples, &got);
  *n_samples = (size_t)got;
  return rc;
}

int ldvbtx_push(ldvbtx_handle *h, const uint8_t *ts_host, size_t n_packets, float *iq_host, size_t cap_samples,
                size_t *n_samples) {
  if (!h || !n_samples || (!ts_host && n_packets) || !iq_host) return LDVB_EINVAL;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return tx_fail(h, LDVB_ECUDA, "cudaSetDevice");
  if (n_packets > h->cfg.max_packets) return tx_fail(h, LDVB_EOVERFLOW, "n_packets exceeds max_packets");
  if (!h->d_ts_in.p) {
    TCK(h->d_ts_in.alloc(h->cfg.max_packets * 188));
    TCK(h->d_out.alloc(h->max_out * 8));
  }
  const TxPlan p = tx_plan(h, n_packets);
  if (p.n_out > cap_samples) return tx_fail(h, LDVB_EOVERFLOW, "output buffer too small (see ldvbtx_max_samples)");
  if (n_packets) TCK(cudaMemcpyAsync(h->d_ts_in.p, ts_host, n_packets * 188, cudaMemcpyHostToDevice, h->st));
  uint64_t got = 0;
  const int rc = tx_process(h, h->d_ts_in.as<uint8_t>(), n_packets, h->d_out.as<float2>(), h->max_out, &got);
  if (rc) return rc;
  if (got) TCK(cudaMemcpyAsync(iq_host, h->d_out.p, got * 8, cudaMemcpyDeviceToHost, h->st));
  TCK(cudaStreamSynchronize(h->st));
  *n_samples = (size_t)got;
  return LDVB_OK;
}

int ldvbtx_tsgen_device(ldvbtx_handle *h, uint64_t first, size_t n_packets, uint8_t *ts_dev) {
  if (!h || (!ts_dev && n_packets)) return LDVB_EINVAL;
  if ((reinterpret_cast<uintptr_t>(ts_dev) & 3u) != 0) return tx_fail(h, LDVB_EINVAL, "ts_dev must be 4-byte aligned");
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return tx_fail(h, LDVB_ECUDA, "cudaSetDevice");
  if (!n_packets) return LDVB_OK;
  k_tx_tsgen<<<(unsigned)((n_packets * 47 + 255) / 256), 256, 0, h->st>>>(first, n_packets, ts_dev);
  TCK(cudaGetLastError());
  return LDVB_OK;
}

int ldvbtx_tap(ldvbtx_handle *h, int which, void *dst, size_t cap, size_t *n_bytes) {
  if (!h || !n_bytes) return LDVB_EINVAL;
  if (!h->cfg.keep_taps) return tx_fail(h, LDVB_ESTATE, "handle was created without keep_taps");
  const DevBuf *b; size_t n;
  switch (which) {
    case LDVBTX_TAP_RSPACKETS: b = &h->t_rs; n = h->tap_rs_packets * 204; break;
    case LDVBTX_TAP_MPEGBYTES: b = &h->t_mb; n = h->tap_mb_bytes; break;
    case LDVBTX_TAP_SYMBOLS: b = &h->t_sym; n = h->tap_symbols; break;
    default: return LDVB_EINVAL;
  }
  *n_bytes = n;
  if (!dst) return LDVB_OK;
  if (cap < n) return LDVB_EOVERFLOW;
  if (n) TCK(cudaMemcpyAsync(dst, b->p, n, cudaMemcpyDeviceToHost, h->st));
  TCK(cudaStreamSynchronize(h->st));
  return LDVB_OK;
}

int ldvbtx_taps(ldvbtx_handle *h, float *dst, size_t cap, size_t *n) {
  if (!h || !n) return LDVB_EINVAL;
  *n = h->taps.size();
  if (dst) {
    if (cap < h->taps.size()) return LDVB_EOVERFLOW;
    memcpy(dst, h->taps.data(), h->taps.size() * 4);
  }
  return LDVB_OK;
}

int ldvbtx_fir_resampler_cf32(int device, const float *x, size_t n_in, const float *taps, uint32_t ntaps, uint32_t interp,
                              float *y, size_t cap_out, size_t *n_out) {
  if (!x || !taps || !y || !n_out || !ntaps || !interp || (size_t)ntaps * 8 > 48 * 1024) return LDVB_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) return LDVB_ENODEV;
  *n_out = 0;
  if (n_in < ntaps || n_in * interp < ntaps) return LDVB_OK;              // dsp.h:307, 320
  const uint64_t steps = (n_in * (uint64_t)interp - ntaps) / interp;      // dsp.h:321
  const uint64_t nout = steps * interp;
  if (nout > cap_out) return LDVB_EOVERFLOW;
  if (!nout) return LDVB_OK;
  if (cudaSetDevice(device) != cudaSuccess) return LDVB_ENODEV;
  float2 *dx = nullptr, *dt = nullptr, *dy = nullptr;
  int rc = LDVB_OK;
  if (cudaMalloc(&dx, n_in * 8) != cudaSuccess || cudaMalloc(&dt, (size_t)ntaps * 8) != cudaSuccess ||
      cudaMalloc(&dy, nout * 8) != cudaSuccess) rc = LDVB_ENOMEM;
  if (!rc) {
    cudaMemcpy(dx, x, n_in * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dt, taps, (size_t)ntaps * 8, cudaMemcpyHostToDevice);
    TxResampleArgs a;
    a.sym = nullptr; a.x = dx; a.points = nullptr; a.taps = dt;
    a.ncoeffs = (int)ntaps; a.interp = (int)interp; a.decim = 1; a.latency = (int)((ntaps + interp) / interp);
    a.n0 = 0; a.m0 = 0; a.count = nout; a.out = dy;
    k_tx_resample<false><<<(unsigned)((nout + 255) / 256), 256, (size_t)ntaps * 8>>>(a);
    if (cudaGetLastError() != cudaSuccess || cudaMemcpy(y, dy, nout * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = LDVB_ECUDA;
    else *n_out = (size_t)nout;
  }
  cudaFree(dx); cudaFree(dt); cudaFree(dy);
  return rc;
}

}  // extern "C"
