// k_rx.cu -- K3: constellation receiver (symbol timing + carrier PLL + AGC +
// soft slicer), the serial core of the chain.
//
// Replaces cstln_receiver<f32>::run() (sdr.h:772-915) with its samplers
// (nearest/linear: sdr.h:589-629) and the constellation look-up
// (cstln_lut<256>::lookup, sdr.h:470-486).  Every statement of the per-sample
// loop is restated with explicit round-to-nearest mul/add (no FMA), truncating
// float->int conversions and the host-built tables, so that a span started from
// the true carry state reproduces every softsymbol field bit for bit.
//
// One thread = one time span of 128-sample chunks:
//   EXACT mode: a single span covers the batch (1 thread, the reference order).
//   FAST  mode: span j owns chunks [j*S, (j+1)*S); it starts W chunks early from
//               a warm-up state, runs kRxVerifyChunks past its end, and logs
//               (time, hard symbol) on both sides of each seam so that
//               k_rx_stitch can align, de-rotate and VERIFY neighbouring spans.
// The IQ stream is read with 16-byte read-only loads (two samples per load);
// the two tables (trig16 512 KB, cstln 512 KB) are read through the L1/L2
// read-only path: accesses cluster around the current phase and the
// constellation points, so they stay L1-resident.
#include "common.cuh"
#include "kernels.h"

namespace ldvb {

namespace {

struct RxRun {
  float mu, phase, freqw, est_insp, agc_gain, est_sp, est_ep;
  float h0pr, h0pi, h0cr, h0ci, h1pr, h1pi, h1cr, h1ci, h2pr, h2pi, h2cr, h2ci;
  float samp_freqw, freq_tap;
  uint32_t meas_count;
};

__device__ __forceinline__ void load_state(RxRun &r, const RxState &s) {
  r.mu = s.mu; r.phase = s.phase; r.freqw = s.freqw; r.est_insp = s.est_insp;
  r.agc_gain = s.agc_gain; r.est_sp = s.est_sp; r.est_ep = s.est_ep;
  r.h0pr = s.hist[0]; r.h0pi = s.hist[1]; r.h0cr = s.hist[2]; r.h0ci = s.hist[3];
  r.h1pr = s.hist[4]; r.h1pi = s.hist[5]; r.h1cr = s.hist[6]; r.h1ci = s.hist[7];
  r.h2pr = s.hist[8]; r.h2pi = s.hist[9]; r.h2cr = s.hist[10]; r.h2ci = s.hist[11];
  r.samp_freqw = s.samp_freqw; r.freq_tap = s.freq_tap; r.meas_count = s.meas_count;
}

__device__ __forceinline__ void store_state(RxState &s, const RxRun &r) {
  s.mu = r.mu; s.phase = r.phase; s.freqw = r.freqw; s.est_insp = r.est_insp;
  s.agc_gain = r.agc_gain; s.est_sp = r.est_sp; s.est_ep = r.est_ep;
  s.hist[0] = r.h0pr; s.hist[1] = r.h0pi; s.hist[2] = r.h0cr; s.hist[3] = r.h0ci;
  s.hist[4] = r.h1pr; s.hist[5] = r.h1pi; s.hist[6] = r.h1cr; s.hist[7] = r.h1ci;
  s.hist[8] = r.h2pr; s.hist[9] = r.h2pi; s.hist[10] = r.h2cr; s.hist[11] = r.h2ci;
  s.samp_freqw = r.samp_freqw; s.freq_tap = r.freq_tap; s.meas_count = r.meas_count;
  s.rrc_update_phase = 0; s.pad = 0;
}

// trig16::expi(float) (math.h:104-110): index = (uint16)(int16)(int32)a.
__device__ __forceinline__ float2 expi(const float2 *__restrict__ trig, float a) {
  return __ldg(trig + ((uint32_t)f2i_trunc(a) & 0xffffu));
}

// One 128-sample chunk of cstln_receiver::run().  `xs` points at the chunk's
// first sample.  Emits symbols through `emit(softsymbol_word, n_local, mu)`.
template <int SAMPLER, class Emit>
__device__ __forceinline__ void rx_chunk(const RxParams &p, RxRun &r, const float2 *__restrict__ xs,
                                         Emit &&emit, float2 *sampled, uint32_t *sampled_flag) {
  if (SAMPLER == 1) r.samp_freqw = r.freqw;  // linear_sampler::update_freq (sdr.h:620)
  float sg_re = 0.f, sg_im = 0.f, s_re = 0.f, s_im = 0.f;
  int have_point = 0;
  float cp_re = 0.f, cp_im = 0.f;

  // Samples are consumed two at a time from 16-byte loads; `cur` is pin[0],
  // `nxt` is pin[1] (read-ahead of the linear sampler).
  const float4 *x4 = reinterpret_cast<const float4 *>(xs);
  float4 w = __ldg(x4);
  float2 cur = make_float2(w.x, w.y);
  float2 nxt = make_float2(w.z, w.w);
#pragma unroll 2
  for (int n = 0; n < kRxChunk; ++n) {
    // Fetch the sample after `nxt` every second step.
    float2 nxt2;
    if ((n & 1) == 0) {
      w = __ldg(x4 + (n >> 1) + 1);
      nxt2 = make_float2(w.x, w.y);
    } else {
      nxt2 = make_float2(w.z, w.w);
    }
    if (r.mu < 1.0f) {
      // --- sampler (sdr.h:595-597, 609-618)
      float2 e0 = expi(p.trig, -r.phase);
      float2 s0 = cmul(cur, e0);
      if (SAMPLER == 1) {
        float2 e1 = expi(p.trig, -fadd(r.phase, r.samp_freqw));
        float2 s1 = cmul(nxt, e1);
        float a = fsub(1.0f, r.mu);
        sg_re = fadd(fmul(s0.x, a), fmul(s1.x, r.mu));
        sg_im = fadd(fmul(s0.y, a), fmul(s1.y, r.mu));
      } else {
        sg_re = s0.x; sg_im = s0.y;
      }
      s_re = fmul(sg_re, r.agc_gain);
      s_im = fmul(sg_im, r.agc_gain);
      // --- constellation look-up (sdr.h:470-486)
      float I = s_re, Q = s_im;
      while (I < -128.f || I > 127.f || Q < -128.f || Q > 127.f) { I = fmul(I, 0.5f); Q = fmul(Q, 0.5f); }
      const uint32_t ci = ((uint32_t)f2i_trunc(I) & 0xffu) * 256u + ((uint32_t)f2i_trunc(Q) & 0xffu);
      const uint2 cellw = __ldg(reinterpret_cast<const uint2 *>(p.cstln) + ci);
      const int cost = (int)(short)(cellw.x & 0xffffu);
      const int symbol = (int)(cellw.x >> 16) & 0xff;
      const int pe = (int)(short)(cellw.y & 0xffffu);
      emit(((uint32_t)cost & 0xffffu) | ((uint32_t)symbol << 16), n, r.mu);
      // --- PLL (sdr.h:814-816)
      const float pef = (float)pe;
      r.phase = fadd(r.phase, fmul(pef, p.freq_alpha));
      r.freqw = fadd(r.freqw, fmul(pef, p.freq_beta));
      // --- modified Mueller & Muller (sdr.h:818-840)
      r.h2pr = r.h1pr; r.h2pi = r.h1pi; r.h2cr = r.h1cr; r.h2ci = r.h1ci;
      r.h1pr = r.h0pr; r.h1pi = r.h0pi; r.h1cr = r.h0cr; r.h1ci = r.h0ci;
      r.h0pr = s_re; r.h0pi = s_im;
      cp_re = (float)p.sym_re[symbol]; cp_im = (float)p.sym_im[symbol];
      have_point = 1;
      r.h0cr = cp_re; r.h0ci = cp_im;
      const float t1 = fadd(fmul(fsub(r.h0pr, r.h2pr), r.h1cr), fmul(fsub(r.h0pi, r.h2pi), r.h1ci));
      const float t2 = fadd(fmul(fsub(r.h0cr, r.h2cr), r.h1pr), fmul(fsub(r.h0ci, r.h2ci), r.h1pi));
      const float muerr = fsub(t1, t2);
      float mucorr = fmul(muerr, p.gain_mu);
      if (mucorr < -0.1f) mucorr = -0.1f;
      if (mucorr > 0.1f) mucorr = 0.1f;
      r.mu = fadd(r.mu, mucorr);
      r.mu = fadd(r.mu, p.omega);
    }
    cur = nxt; nxt = nxt2;
    r.mu = fsub(r.mu, 1.0f);
    r.phase = fadd(r.phase, r.freqw);
  }

  r.phase = fmodf(r.phase, 65536.0f);  // sdr.h:855 (fmodf is exact)

  if (have_point) {
    if (sampled) { *sampled = make_float2(s_re, s_im); *sampled_flag = 1; }
    // AGC (sdr.h:863-869)
    const float insp = fadd(fmul(sg_re, sg_re), fmul(sg_im, sg_im));
    const float omk = fsub(1.0f, p.kest);
    r.est_insp = fadd(fmul(insp, p.kest), fmul(r.est_insp, omk));
    if (r.est_insp != 0.0f) r.agc_gain = __fdiv_rn(75.0f, __fsqrt_rn(r.est_insp));
    // SS / MER estimators (sdr.h:871-888)
    const float ev_re = fsub(s_re, cp_re), ev_im = fsub(s_im, cp_im);
    float sig_power, ev_power;
    if (p.nsymbols == 2) {
      // (float)((int + int) * 0.707) and (float)((float + float) * 0.707): double products
      const float sig_real = (float)__dmul_rn((double)(int)(cp_re + cp_im), 0.707);
      const float ev_real = (float)__dmul_rn((double)fadd(ev_re, ev_im), 0.707);
      sig_power = fmul(sig_real, sig_real);
      ev_power = fmul(ev_real, ev_real);
    } else {
      const int ire = (int)cp_re, iim = (int)cp_im;
      sig_power = (float)(ire * ire + iim * iim);
      ev_power = fadd(fmul(ev_re, ev_re), fmul(ev_im, ev_im));
    }
    r.est_sp = fadd(fmul(sig_power, p.kest), fmul(r.est_sp, omk));
    r.est_ep = fadd(fmul(ev_power, p.kest), fmul(r.est_ep, omk));
  } else if (sampled_flag) {
    *sampled_flag = 0;
  }

  if (!p.allow_drift) {  // sdr.h:895-898
    if (r.freqw < p.min_freqw || r.freqw > p.max_freqw)
      r.freqw = __fdiv_rn(fadd(p.max_freqw, p.min_freqw), 2.0f);
  }
  r.freq_tap = __fdiv_rn(r.freqw, 65536.0f);  // sdr.h:917-919
}

template <int SAMPLER>
__device__ void rx_span(const RxArgs &a, uint32_t span, const RxState *forced) {
  const RxParams &p = a.p;
  const bool exact_start = (span == 0) || (forced != nullptr);
  const uint64_t own_begin = (uint64_t)span * a.span_chunks;
  uint64_t own_end = own_begin + a.span_chunks;
  if (own_end > a.nchunks) own_end = a.nchunks;
  const bool last = (own_end >= a.nchunks);
  uint64_t run_begin = own_begin;
  RxRun r;
  if (forced) load_state(r, *forced);
  else load_state(r, *a.state_in);
  if (!exact_start) {
    // Warm-up: start W chunks early from the carried loop state with the
    // timing / phase registers cleared.
    run_begin = (own_begin > a.warm_chunks) ? own_begin - a.warm_chunks : 0;
    r.mu = 0.f; r.phase = 0.f;
    r.h0pr = r.h0pi = r.h0cr = r.h0ci = 0.f;
    r.h1pr = r.h1pi = r.h1cr = r.h1ci = 0.f;
    r.h2pr = r.h2pi = r.h2cr = r.h2ci = 0.f;
    // meas_count is a pure function of the position
    uint64_t mc = ((uint64_t)a.state_in->meas_count + run_begin * (uint64_t)kRxChunk) % p.meas_decimation;
    r.meas_count = (uint32_t)mc;
  }
  uint64_t run_end = own_end;
  if (!last) {
    run_end = own_end + kRxVerifyChunks;
    if (run_end > a.nchunks) run_end = a.nchunks;
  }

  uint32_t *out = a.sym_out + (size_t)span * a.span_cap;
  RxSeamSym *hlog = a.head_log ? a.head_log + (size_t)span * kRxSeamLog : nullptr;
  RxSeamSym *tlog = a.tail_log ? a.tail_log + (size_t)span * kRxSeamLog : nullptr;
  uint32_t n_out = 0, n_tail = 0, n_head = 0;
  const uint32_t cap = a.span_cap;

  for (uint64_t c = run_begin; c < run_end; ++c) {
    const float2 *xs = a.x + c * kRxChunk;
    const int phase_of_run = (c < own_begin) ? 0 : (c < own_end ? 1 : 2);
    const float t_head = (float)((double)(c - own_begin) * kRxChunk);  // chunk offset from the head seam
    const float t_tail = (float)((double)(c - own_end) * kRxChunk);
    const bool log_head = hlog && span > 0 && phase_of_run == 1 && (c - own_begin) < kRxVerifyChunks;
    auto emit = [&](uint32_t word, int n, float mu) {
      if (phase_of_run == 1) {
        if (n_out < cap) out[n_out] = word;
        ++n_out;
        if (log_head && n_head < kRxSeamLog) {
          hlog[n_head].t = t_head + (float)n + mu;
          hlog[n_head].sym = word >> 16;
          ++n_head;
        }
      } else if (phase_of_run == 2) {
        // Verification overlap: stored right after the owned symbols so that the
        // stitcher can extend this span by one symbol when needed.
        if (n_out + n_tail < cap) out[n_out + n_tail] = word;
        if (tlog && n_tail < kRxSeamLog) {
          tlog[n_tail].t = t_tail + (float)n + mu;
          tlog[n_tail].sym = word >> 16;
        }
        ++n_tail;
      }
    };
    float2 *smp = nullptr; uint32_t *smpf = nullptr;
    if (a.sampled && phase_of_run == 1) { smp = a.sampled + c; smpf = a.sampled_flag + c; }
    rx_chunk<SAMPLER>(p, r, xs, emit, smp, smpf);

    // Measurements (sdr.h:904-913)
    r.meas_count += kRxChunk;
    while (r.meas_count >= p.meas_decimation) {
      r.meas_count -= p.meas_decimation;
      if (a.meas && phase_of_run == 1) {
        uint32_t k = atomicAdd(a.meas_count, 1u);
        if (k < a.max_meas) {
          float *m = a.meas + 4 * (size_t)k;
          m[0] = (float)c;
          m[1] = r.freq_tap;
          m[2] = __fsqrt_rn(r.est_insp);
          // 10*logf(sp/ep)/logf(10): evaluated on the host from (sp, ep) when exactness
          // matters; here the device logf is used for telemetry only.
          m[3] = (r.est_ep != 0.0f) ? 10.0f * logf(__fdiv_rn(r.est_sp, r.est_ep)) / logf(10.0f) : 0.0f;
        }
      }
    }
    if (c + 1 == own_end) store_state(a.state_end[span], r);
  }
  if (own_end <= run_begin) store_state(a.state_end[span], r);
  RxSpanInfo inf;
  inf.n_out = n_out; inf.n_tail = n_tail; inf.n_head_logged = n_head; inf.pad = 0;
  a.info[span] = inf;
}

__global__ void __launch_bounds__(32)
k_rx(RxArgs a, int only_span, const RxState *forced) {
  uint32_t span = (only_span >= 0) ? (uint32_t)only_span : blockIdx.x * blockDim.x + threadIdx.x;
  if (span >= a.nspans) return;
  if (only_span >= 0 && (blockIdx.x != 0 || threadIdx.x != 0)) return;
  if (a.p.sampler == 0) rx_span<0>(a, span, forced);
  else rx_span<1>(a, span, forced);
}

// ---------------------------------------------------------------- seam stitching

__global__ void k_rx_stitch(RxStitchArgs a, int only_seam) {
  uint32_t j = (only_seam >= 0) ? (uint32_t)only_seam : blockIdx.x * blockDim.x + threadIdx.x;
  if (j + 1 >= a.nspans) return;
  if (only_seam >= 0 && (blockIdx.x != 0 || threadIdx.x != 0)) return;
  const RxSeamSym *tail = a.tail_log + (size_t)j * kRxSeamLog;
  const RxSeamSym *head = a.head_log + (size_t)(j + 1) * kRxSeamLog;
  uint32_t nt = min(a.info[j].n_tail, (uint32_t)kRxSeamLog);
  uint32_t nh = min(a.info[j + 1].n_head_logged, (uint32_t)kRxSeamLog);
  RxSeam s;
  s.ok = 0; s.rot = 0; s.extend_prev = 0; s.skip_next = 0; s.compared = 0; s.mismatches = 0;
  if (nt >= 8 && nh >= 8) {
    // Align on symbol time: tail[it0 + i] <-> head[ih0 + i].
    const float half = 0.5f * a.omega;
    int it0 = 0, ih0 = 0;
    const float d = head[0].t - tail[0].t;
    if (d > half) { it0 = 1; s.extend_prev = 1; }        // next span missed the first symbol
    else if (d < -half) { ih0 = 1; s.skip_next = 1; }    // next span repeats the previous span's last symbol
    const int n = (int)min(nt - it0, nh - ih0);
    int best_rot = -1, best_mis = 1 << 30;
    for (int rot = 0; rot < a.nrot; ++rot) {
      const uint8_t *perm = a.rot_perm + rot * a.nsymbols;
      int mis = 0;
      for (int i = 0; i < n; ++i)
        if (perm[head[ih0 + i].sym] != tail[it0 + i].sym) ++mis;
      if (mis < best_mis) { best_mis = mis; best_rot = rot; }
    }
    bool time_ok = true;
    for (int i = 0; i < n; ++i)
      if (fabsf(head[ih0 + i].t - tail[it0 + i].t) > 0.25f * a.omega) time_ok = false;
    s.rot = best_rot; s.compared = n; s.mismatches = best_mis;
    s.ok = (time_ok && best_mis == 0 && n >= 8) ? 1 : 0;
  }
  a.seams[j] = s;
}

__global__ void k_rx_compact(RxCompactArgs a, uint64_t total) {
  // One block per (span, slice); threads copy with the span's rotation applied.
  const uint32_t span = blockIdx.y;
  const uint64_t base = a.span_offset[span];
  const uint64_t n = a.span_offset[span + 1] - base;
  const uint32_t *src = a.sym_in + (size_t)span * a.span_cap + a.span_skip[span];
  const uint8_t *perm = a.rot_perm + (size_t)a.span_rot[span] * a.nsymbols;
  const bool identity = (a.span_rot[span] == 0);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t w = src[i];
    if (!identity) {
      uint32_t sym = (w >> 16) & 0xffu;
      w = (w & 0xffffu) | ((uint32_t)perm[sym] << 16);
    }
    a.sym_out[base + i] = w;
  }
}

}  // namespace

cudaError_t launch_rx(const RxArgs &a, int only_span, const RxState *forced, cudaStream_t st) {
  if (a.nspans == 0) return cudaSuccess;
  if (only_span >= 0) {
    k_rx<<<1, 32, 0, st>>>(a, only_span, forced);
  } else {
    const unsigned threads = 32;
    const unsigned blocks = (a.nspans + threads - 1) / threads;
    k_rx<<<blocks, threads, 0, st>>>(a, -1, nullptr);
  }
  return cudaGetLastError();
}

cudaError_t launch_rx_stitch(const RxStitchArgs &a, int only_seam, cudaStream_t st) {
  if (a.nspans < 2) return cudaSuccess;
  if (only_seam >= 0) k_rx_stitch<<<1, 32, 0, st>>>(a, only_seam);
  else k_rx_stitch<<<(a.nspans - 1 + 63) / 64, 64, 0, st>>>(a, -1);
  return cudaGetLastError();
}

cudaError_t launch_rx_compact(const RxCompactArgs &a, uint64_t total, cudaStream_t st) {
  if (a.nspans == 0 || total == 0) return cudaSuccess;
  dim3 grid(8, a.nspans);
  k_rx_compact<<<grid, 256, 0, st>>>(a, total);
  return cudaGetLastError();
}

}  // namespace ldvb
