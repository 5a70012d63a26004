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
// Parallel decomposition: one LANE = one time span of 128-sample chunks.
//   EXACT mode: a single span covers the batch (the reference order).
//   FAST  mode: span j owns chunks [j*S, (j+1)*S); it starts W chunks early from
//               a warm-up state, runs kRxVerifyChunks past its end, and logs
//               (time, hard symbol) on both sides of each seam so that
//               k_rx_stitch can align, de-rotate and VERIFY neighbouring spans.
//
// Data movement: the 32 lanes of a warp walk 32 different places of the IQ
// stream in lock step.  Per tile the warp stages, for every lane, that lane's
// next 32 samples (+2 look-ahead) in a private shared-memory row: 32 rows x
// 272 B, fetched cooperatively with 16-byte asynchronous copies (cp.async.cg,
// SASS LDGSTS; contiguous within a row), double buffered.  The row pitch (304 B =
// 76 words, 12 mod 32) makes the 16-byte row reads of any 8 consecutive lanes
// bank-conflict free.  HBM sees full, aligned 304-byte bursts and the recurrence never waits on a global
// load.  (One TMA bulk copy per row was measured first: ~0.6 us per 272-byte
// request, serialised per SM -- rows this small are below the TMA's grain.)  The two tables (trig16, cstln: 512 KB each) are read
// through the read-only path; their accesses cluster around the current carrier
// phase and the constellation points.
#include "common.cuh"
#include <cstdlib>

#include "kernels.h"

namespace ldvb {

namespace {

// Samples per staged tile: template parameter TILE (8 or 16) of rx_warp / k_rx.
// Row = tile + look-ahead: 2 samples for the nearest/linear samplers (272 B), 6 for the RRC
// sampler (304 B, up to 6 taps).  Both pitches are 16 B mod 128-friendly: any 8 consecutive
// lanes read 16 B each from distinct banks.
template <int SAMPLER, int TILE> struct RowCfg { static constexpr int kBytes = (TILE + (SAMPLER == 2 ? 6 : 2)) * 8; };
// Slicer variants (template parameter SLICER of rx_sample / rx_warp):
//   0  cstln_lut<256>::lookup as a gather into the 512 KB cell table in global memory (any constellation);
//   1  QPSK: symbol and cost follow from the truncated (I, Q) by arithmetic
//      (tests/test_capi_cpu.py::test_qpsk_table_cells_follow_from_arithmetic checks all 65536 cells of the
//      host-built table), and the phase error -- glibc atan2f, so it stays a table -- is read from a 66 KB
//      int16 copy of that column (folded over Q) held in SHARED memory: the constellation look-up leaves the L1/L2 path
//      altogether (round 1, ncu: one lane of every warp-wide gather missed L1, so every symbol waited for L2).
constexpr int kPe16Bytes = 256 * kPeFoldPitch * 2;   // folded over Q (kernels.h): 66 KB
static_assert(kPe16Bytes % 16 == 0, "copied 16 bytes at a time");
#ifndef LDVB_RX_CA
#define LDVB_RX_CA 0
#endif
constexpr bool kRxCa = LDVB_RX_CA != 0;
#ifndef LDVB_RX_STAGES
#define LDVB_RX_STAGES 2
#endif
constexpr int kStages = LDVB_RX_STAGES;   // tiles in flight per lane (kStages - 1 ahead of the one in use)

struct RxRun {
  float mu, phase, freqw, est_insp, agc_gain, est_sp, est_ep;
  float h0pr, h0pi, h0cr, h0ci, h1pr, h1pi, h1cr, h1ci, h2pr, h2pi, h2cr, h2ci;
  float samp_freqw, freq_tap;
  uint32_t meas_count;
  // fir_sampler (sdr.h:635-689): frequency the taps were last shifted for, update throttle
  float rrc_f;
  int rrc_update_phase;
  // per-chunk scratch (sdr.h:795-798)
  float sg_re, sg_im, s_re, s_im, cp_re, cp_im;
  int have_point;
};

__device__ __forceinline__ void load_state(RxRun &r, const RxState &s) {
  r.mu = s.mu; r.phase = s.phase; r.freqw = s.freqw; r.est_insp = s.est_insp;
  r.agc_gain = s.agc_gain; r.est_sp = s.est_sp; r.est_ep = s.est_ep;
  r.h0pr = s.hist[0]; r.h0pi = s.hist[1]; r.h0cr = s.hist[2]; r.h0ci = s.hist[3];
  r.h1pr = s.hist[4]; r.h1pi = s.hist[5]; r.h1cr = s.hist[6]; r.h1ci = s.hist[7];
  r.h2pr = s.hist[8]; r.h2pi = s.hist[9]; r.h2cr = s.hist[10]; r.h2ci = s.hist[11];
  r.samp_freqw = s.samp_freqw; r.freq_tap = s.freq_tap; r.meas_count = s.meas_count;
  r.rrc_f = s.rrc_f; r.rrc_update_phase = s.rrc_update_phase;
}

__device__ __forceinline__ void store_state(RxState &s, const RxRun &r) {
  s.mu = r.mu; s.phase = r.phase; s.freqw = r.freqw; s.est_insp = r.est_insp;
  s.agc_gain = r.agc_gain; s.est_sp = r.est_sp; s.est_ep = r.est_ep;
  s.hist[0] = r.h0pr; s.hist[1] = r.h0pi; s.hist[2] = r.h0cr; s.hist[3] = r.h0ci;
  s.hist[4] = r.h1pr; s.hist[5] = r.h1pi; s.hist[6] = r.h1cr; s.hist[7] = r.h1ci;
  s.hist[8] = r.h2pr; s.hist[9] = r.h2pi; s.hist[10] = r.h2cr; s.hist[11] = r.h2ci;
  s.samp_freqw = r.samp_freqw; s.freq_tap = r.freq_tap; s.meas_count = r.meas_count;
  s.rrc_update_phase = r.rrc_update_phase; s.rrc_f = r.rrc_f;
}

// Signed 16-bit load from a shared-space byte address (no generic-address conversion on the dependent path).
__device__ __forceinline__ int lds_s16(uint32_t addr) {
  short v;
  asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(addr));
  return (int)v;
}

// trig16::expi(float) (math.h:104-110): index = (uint16)(int16)(int32)a.
#ifndef LDVB_RX_TRIG_EVICT_LAST
#define LDVB_RX_TRIG_EVICT_LAST 1
#endif
__device__ __forceinline__ float2 expi(const float2 *__restrict__ trig, float a) {
  const float2 *p = trig + ((uint32_t)f2i_trunc(a) & 0xffffu);
#if LDVB_RX_TRIG_EVICT_LAST
  // The table lines around the current carrier phase are the only data of this kernel that L1 should keep.
  float2 v;
  asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

// One input sample of cstln_receiver::run()'s inner loop (sdr.h:800-847).
// cur = pin[0], nxt = pin[1].  Returns true and fills `word` when a symbol is emitted.
template <int SAMPLER, int SLICER>
__device__ __forceinline__ bool rx_sample(const RxParams &p, RxRun &r, float2 cur, float2 nxt,
                                          const float2 *pin, uint32_t s_pe, uint32_t &word, float &mu_emit) {
  bool emitted = false;
  if (r.mu < 1.0f) {
    // --- sampler (sdr.h:595-597, 609-618, 647-665)
    const float2 e0 = expi(p.trig, -r.phase);
    const float2 s0 = cmul(cur, e0);
    if (SAMPLER == 2) {
      // fir_sampler::interp: acc = sum_j shifted[off + j*sub] * pin[j], then derotate.
      // shifted[i] = expi(-f*(i - n/2)) * coeffs[i] (do_update_freq, sdr.h:678-682) is
      // re-evaluated from the saved f instead of being kept as a per-lane table.
      const int off = f2i_trunc(fmul(fsub(1.0f, r.mu), (float)p.rrc_sub));
      float acc_re = 0.f, acc_im = 0.f;
      int j = 0;
      for (int ci = off; ci < p.rrc_n; ci += p.rrc_sub, ++j) {
        const float2 e = expi(p.trig, fmul(-r.rrc_f, (float)(ci - p.rrc_n / 2)));
        const float cf = __ldg(p.rrc_coeffs + ci);
        const float2 pr = cmul(make_float2(fmul(e.x, cf), fmul(e.y, cf)), pin[j]);
        acc_re = fadd(acc_re, pr.x);
        acc_im = fadd(acc_im, pr.y);
      }
      const float2 sg = cmul(e0, make_float2(acc_re, acc_im));
      r.sg_re = sg.x; r.sg_im = sg.y;
    } else if (SAMPLER == 1) {
      const float2 e1 = expi(p.trig, -fadd(r.phase, r.samp_freqw));
      const float2 s1 = cmul(nxt, e1);
      const float a = fsub(1.0f, r.mu);
      r.sg_re = fadd(fmul(s0.x, a), fmul(s1.x, r.mu));
      r.sg_im = fadd(fmul(s0.y, a), fmul(s1.y, r.mu));
    } else {
      r.sg_re = s0.x; r.sg_im = s0.y;
    }
    r.s_re = fmul(r.sg_re, r.agc_gain);
    r.s_im = fmul(r.sg_im, r.agc_gain);
    // --- constellation look-up (sdr.h:470-486)
    float I = r.s_re, Q = r.s_im;
    // The truncations are issued before the range test resolves (the halving loop is rare): the recurrence
    // is one long dependent chain, so the test must not sit in front of the conversions.
    int Ii, Qi;
    asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(Ii) : "f"(I));   // (volatile: keeps them in front of the test)
    asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(Qi) : "f"(Q));
    if (I < -128.f || I > 127.f || Q < -128.f || Q > 127.f) {
      do { I = fmul(I, 0.5f); Q = fmul(Q, 0.5f); } while (I < -128.f || I > 127.f || Q < -128.f || Q > 127.f);
      Ii = f2i_trunc(I); Qi = f2i_trunc(Q);
    }
    int symbol, pe;
    if (SLICER == 1) {
      // QPSK: the cell's symbol and cost follow from the truncated (I, Q); only the phase error is looked up
      // (shared memory: a 2-byte gather, ~3 bank wavefronts for 32 random lanes).
      const int aI = abs(Ii), aQ = abs(Qi);
      // folded column: [I & 0xff][|Q|], pe(I, -Q) = -pe(I, Q)
      pe = lds_s16(s_pe + 2u * (((uint32_t)Ii & 0xffu) * (uint32_t)kPeFoldPitch + (uint32_t)aQ));
      if (Qi < 0) pe = -pe;
      symbol = ((Ii < 0) ? 2 : 0) | ((Qi < 0) ? 1 : 0);
      const int d1 = (aI - 53) * (aI - 53) + (aQ - 53) * (aQ - 53);
      const int d2 = d1 + 212 * min(aI, aQ);
      const int cost = min(d1, 32767) - min(d2, 32767);
      word = ((uint32_t)cost & 0xffffu) | ((uint32_t)symbol << 16);
    } else {
      const uint32_t ci = __byte_perm((uint32_t)Qi, (uint32_t)Ii, 0x0040) & 0xffffu;
      const uint2 cellw = __ldg(reinterpret_cast<const uint2 *>(p.cstln) + ci);
      symbol = (int)(cellw.x >> 16) & 0xff;
      pe = (int)(short)(cellw.y & 0xffffu);
      word = (cellw.x & 0xffffu) | ((uint32_t)symbol << 16);
    }
    mu_emit = r.mu;
    emitted = true;
    // --- PLL (sdr.h:814-816)
    const float pef = (float)pe;
    r.phase = fadd(r.phase, fmul(pef, p.freq_alpha));
    r.freqw = fadd(r.freqw, fmul(pef, p.freq_beta));
    // --- modified Mueller & Muller (sdr.h:818-840)
    r.h2pr = r.h1pr; r.h2pi = r.h1pi; r.h2cr = r.h1cr; r.h2ci = r.h1ci;
    r.h1pr = r.h0pr; r.h1pi = r.h0pi; r.h1cr = r.h0cr; r.h1ci = r.h0ci;
    r.h0pr = r.s_re; r.h0pi = r.s_im;
    if (SLICER == 1) { r.cp_re = (symbol & 2) ? -53.0f : 53.0f; r.cp_im = (symbol & 1) ? -53.0f : 53.0f; }
    else { r.cp_re = (float)p.sym_re[symbol]; r.cp_im = (float)p.sym_im[symbol]; }
    r.have_point = 1;
    r.h0cr = r.cp_re; r.h0ci = r.cp_im;
    const float t1 = fadd(fmul(fsub(r.h0pr, r.h2pr), r.h1cr), fmul(fsub(r.h0pi, r.h2pi), r.h1ci));
    const float t2 = fadd(fmul(fsub(r.h0cr, r.h2cr), r.h1pr), fmul(fsub(r.h0ci, r.h2ci), r.h1pi));
    const float muerr = fsub(t1, t2);
    float mucorr = fmul(muerr, p.gain_mu);
    if (mucorr < -0.1f) mucorr = -0.1f;
    if (mucorr > 0.1f) mucorr = 0.1f;
    r.mu = fadd(r.mu, mucorr);
    r.mu = fadd(r.mu, p.omega);
  }
  r.mu = fsub(r.mu, 1.0f);
  r.phase = fadd(r.phase, r.freqw);
  return emitted;
}

// One input sample of fast_qpsk_receiver::run()'s inner loop (sdr.h:1024-1111).  Integer
// arithmetic as in the reference (u_angle = uint16_t, signed long = 64 bits); the state lives in
// RxRun's float fields as exactly representable integers.  cur/nxt are the u8 samples as the
// front end converted them (value - 128, exact).
__device__ __forceinline__ bool rx_sample_hs(const RxParams &p, RxRun &r, float2 cur, float2 nxt, uint32_t &word,
                                             float &mu_emit) {
  bool emitted = false;
  int phase = (int)r.phase;
  long long freqw = (long long)r.freqw;
  if (r.mu < 1.0f) {
    const uint32_t i0 = (uint32_t)((int)cur.x + 128) * 256u + (uint32_t)((int)cur.y + 128);
    const uint32_t i1 = (uint32_t)((int)nxt.x + 128) * 256u + (uint32_t)((int)nxt.y + 128);
    const uint32_t p0 = __ldg(p.hs_polar + i0), p1 = __ldg(p.hs_polar + i1);
    const uint32_t a0 = (((p0 & 0xffffu) - (uint32_t)phase) & 0xffffu) >> 8;                        // sdr.h:1040
    const uint32_t a1 = ((uint32_t)((long long)(p1 & 0xffffu) - ((long long)phase + freqw)) & 0xffffu) >> 8;
    const uint32_t r0 = __ldg(p.hs_rect + a0 * 256u + ((p0 >> 16) >> 1));
    const uint32_t r1 = __ldg(p.hs_rect + a1 * 256u + ((p1 >> 16) >> 1));
    const int p0re = (int)(r0 & 0xffu), p0im = (int)(r0 >> 8), p1re = (int)(r1 & 0xffu), p1im = (int)(r1 >> 8);
    // s.re = (int)(p0r->re + (p1r->re - p0r->re)*mu), stored in a u8 (sdr.h:1045-1046)
    const uint32_t sre = (uint32_t)f2i_trunc(fadd((float)p0re, fmul((float)(p1re - p0re), r.mu))) & 0xffu;
    const uint32_t sim = (uint32_t)f2i_trunc(fadd((float)p0im, fmul((float)(p1im - p0im), r.mu))) & 0xffu;
    const uint32_t symbol_arg = __ldg(p.hs_polar + sre * 256u + sim) & 0xffffu;
    const uint32_t q2s = 0x1320u;                                 // quadrant_to_symbol {0,2,3,1}, one nibble each
    const uint32_t sym = (q2s >> (4 * (symbol_arg >> 14))) & 0xfu;
    word = sym << 16;
    mu_emit = r.mu;
    emitted = true;
    const long long pe = (long long)(int)(symbol_arg & 16383u) - 8192;                               // sdr.h:1063
    phase = (int)(((long long)phase + ((pe * 2621 + 32768) >> 16)) & 0xffff);                        // freq_alpha = 0.04*65536
    freqw += (pe * p.hs_freq_beta + 32768 * 256) >> 24;
    r.h2pr = r.h1pr; r.h2pi = r.h1pi; r.h2cr = r.h1cr; r.h2ci = r.h1ci;
    r.h1pr = r.h0pr; r.h1pi = r.h0pi; r.h1cr = r.h0cr; r.h1ci = r.h0ci;
    const uint32_t cp = __ldg(p.hs_sincos + (((symbol_arg & 49152u) + 8192u) & 0xffffu));
    r.h0pr = (float)sre; r.h0pi = (float)sim; r.h0cr = (float)(cp & 0xffu); r.h0ci = (float)(cp >> 8);
    const int h0pr = (int)sre, h0pi = (int)sim, h0cr = (int)(cp & 0xffu), h0ci = (int)(cp >> 8);
    const int h1pr = (int)r.h1pr, h1pi = (int)r.h1pi, h1cr = (int)r.h1cr, h1ci = (int)r.h1ci;
    const int h2pr = (int)r.h2pr, h2pi = (int)r.h2pi, h2cr = (int)r.h2cr, h2ci = (int)r.h2ci;
    const int muerr = ((int)(signed char)(h0pr - h2pr) * (h1cr - 128) + (int)(signed char)(h0pi - h2pi) * (h1ci - 128)) -
                      ((int)(signed char)(h0cr - h2cr) * (h1pr - 128) + (int)(signed char)(h0ci - h2ci) * (h1pi - 128));
    float mucorr = fmul((float)muerr, p.gain_mu);
    if (mucorr < -0.1f) mucorr = -0.1f;
    if (mucorr > 0.1f) mucorr = 0.1f;
    r.mu = fadd(r.mu, mucorr);
    r.mu = fadd(r.mu, p.omega);
  }
  r.mu = fsub(r.mu, 1.0f);
  phase = (int)(((long long)phase + freqw) & 0xffff);
  r.phase = (float)phase;
  r.freqw = (float)freqw;
  return emitted;
}

__device__ __forceinline__ void rx_chunk_begin(const RxParams &p, RxRun &r, int sampler) {
  if (sampler == 1) r.samp_freqw = r.freqw;  // linear_sampler::update_freq (sdr.h:620)
  if (sampler == 2) {                        // fir_sampler::update_freq (sdr.h:667-675)
    r.rrc_update_phase -= kRxChunk;
    if (r.rrc_update_phase <= 0) {
      r.rrc_update_phase = p.rrc_n * 16;
      r.rrc_f = __fdiv_rn(r.freqw, (float)p.rrc_sub);
    }
  }
  r.have_point = 0;
}

// End-of-chunk bookkeeping of fast_qpsk_receiver::run() (sdr.h:1122-1125): integer limits.
__device__ __forceinline__ void rx_chunk_end_hs(const RxParams &p, RxRun &r) {
  if (!p.allow_drift) {
    const long long f = (long long)r.freqw, lo = (long long)p.min_freqw, hi = (long long)p.max_freqw;
    if (f < lo || f > hi) r.freqw = (float)((hi + lo) / 2);
  }
  r.freq_tap = __fdiv_rn(r.freqw, 65536.0f);   // what freq_out reports (sdr.h:1133)
}

// End-of-chunk bookkeeping of cstln_receiver::run() (sdr.h:849-902).
__device__ __forceinline__ void rx_chunk_end(const RxParams &p, RxRun &r) {
  r.phase = fmodf(r.phase, 65536.0f);  // sdr.h:855 (fmodf is exact)
  if (r.have_point) {
    // AGC (sdr.h:863-869)
    const float insp = fadd(fmul(r.sg_re, r.sg_re), fmul(r.sg_im, r.sg_im));
    const float omk = fsub(1.0f, p.kest);
    r.est_insp = fadd(fmul(insp, p.kest), fmul(r.est_insp, omk));
    if (r.est_insp != 0.0f) r.agc_gain = __fdiv_rn(75.0f, __fsqrt_rn(r.est_insp));
    // SS / MER estimators (sdr.h:871-888)
    const float ev_re = fsub(r.s_re, r.cp_re), ev_im = fsub(r.s_im, r.cp_im);
    float sig_power, ev_power;
    if (p.nsymbols == 2) {
      // (float)((int + int) * 0.707) and (float)((float + float) * 0.707): double products
      const float sig_real = (float)__dmul_rn((double)(int)(r.cp_re + r.cp_im), 0.707);
      const float ev_real = (float)__dmul_rn((double)fadd(ev_re, ev_im), 0.707);
      sig_power = fmul(sig_real, sig_real);
      ev_power = fmul(ev_real, ev_real);
    } else {
      const int ire = (int)r.cp_re, iim = (int)r.cp_im;
      sig_power = (float)(ire * ire + iim * iim);
      ev_power = fadd(fmul(ev_re, ev_re), fmul(ev_im, ev_im));
    }
    r.est_sp = fadd(fmul(sig_power, p.kest), fmul(r.est_sp, omk));
    r.est_ep = fadd(fmul(ev_power, p.kest), fmul(r.est_ep, omk));
  }
  if (!p.allow_drift) {  // sdr.h:895-898
    if (r.freqw < p.min_freqw || r.freqw > p.max_freqw)
      r.freqw = __fdiv_rn(fadd(p.max_freqw, p.min_freqw), 2.0f);
  }
  r.freq_tap = __fdiv_rn(r.freqw, 65536.0f);  // sdr.h:917-919
}

// Where the symbols of a lane go.  All counters are per lane.
struct RxEmit {
  uint32_t *out;            // the span's region of sym_out
  RxSeamSym *hlog, *tlog;   // seam logs of the span (or null)
  uint32_t n_out, n_tail, n_head, cap;
  uint32_t w0, w1, w2;      // symbols waiting for the fourth one of their 16-byte group (vec mode)
  bool vec;                 // out is 16-byte aligned and cap a multiple of 4: symbols leave four at a time
};

// 16-byte / 4-byte stores that do not allocate in L1 (the symbols are read by a later kernel; in L1 they would
// only evict the trig16 lines -- ncu, round 2: 139 M store sectors per launch went through L1 next to 322 M
// table sectors, table hit rate 74 %).
__device__ __forceinline__ void st_cg(uint4 *p, uint4 v) { __stcg(p, v); }
__device__ __forceinline__ void st_cg(uint32_t *p, uint32_t v) { __stcg(p, v); }

// Symbol `word` goes to position pos of the span's output (owned symbols, then the verification overlap).
// Each lane writes its own span: one 4-byte store per symbol is a 32-byte sector per lane and instruction;
// in vec mode the lane keeps three words in registers and writes 16 bytes with the fourth.
// (Positions arrive in order, one per call: (w0, w1, w2) is a shift register of the last three words, oldest first.  A
//  branch-free update -- one predicated store and three moves -- instead of a four-way branch on pos & 3 whose arms
//  each wrote another register: 351 instead of 368 SASS instructions per two samples of the owned loop (linear
//  sampler, QPSK slicer; the warm-up loop, same recurrence without emission, has 244), 16 instead of 24 branches, same
//  96 registers; cuobjdump -sass at the end of round 2, checked word for word on the host by tests/emu/emu_rx.cpp.)
__device__ __forceinline__ void emit_word(RxEmit &e, uint32_t pos, uint32_t word) {
  if (!e.vec) { if (pos < e.cap) st_cg(e.out + pos, word); return; }
  if ((pos & 3u) == 3u && pos < e.cap) st_cg(reinterpret_cast<uint4 *>(e.out + (pos - 3u)), make_uint4(e.w0, e.w1, e.w2, word));
  e.w0 = e.w1; e.w1 = e.w2; e.w2 = word;
}
// The words of the last, incomplete group: the last r of (w0, w1, w2).
__device__ __forceinline__ void emit_flush(RxEmit &e) {
  if (!e.vec) return;
  const uint32_t pos = e.n_out + e.n_tail, r = pos & 3u, b = pos - r;
  const uint32_t w[3] = {e.w0, e.w1, e.w2};
  for (uint32_t i = 0; i < r; ++i)
    if (b + i < e.cap) st_cg(e.out + b + i, w[3 - r + i]);
}
__device__ __forceinline__ void emit_init(RxEmit &e) {
  e.n_out = e.n_tail = e.n_head = 0; e.w0 = e.w1 = e.w2 = 0;
  e.vec = ((reinterpret_cast<uintptr_t>(e.out) & 15u) == 0) && ((e.cap & 3u) == 0);
}

// What a lane does with the symbols of the chunk it is walking.
//   kWarm   warm-up in front of the span: loops run, nothing is kept
//   kOwned  owned chunk: symbols are appended to the span's output
//   kHead   first owned chunk(s): as kOwned + (time, hard symbol) into the head log
//   kTail   verification overlap behind the span: symbols stored after the owned ones + tail log
enum { kWarm = 0, kOwned = 1, kHead = 2, kTail = 3 };

// One staged tile (TILE samples + look-ahead, at `rp`) of one lane.  MODE is a template parameter so that the
// per-symbol bookkeeping of the three output flavours is not evaluated symbol by symbol (the recurrence is
// latency bound: every predicate in its way costs issue slots of the only instruction stream the lane has).
template <int SAMPLER, int SLICER, int TILE, int MODE>
__device__ __forceinline__ void rx_tile(const RxParams &p, RxRun &r, const float4 *rp, uint32_t s_pe, RxEmit &e,
                                        float t0 /* time of the tile's first sample relative to the seam */) {
  float4 w = rp[0];
  float2 cur = make_float2(w.x, w.y), nxt = make_float2(w.z, w.w);
  // (unrolled by two -- one 16-byte row read per pair of samples -- and no further: the body is ~140
  //  instructions per sample, and four modes x three samplers of it have to stay in the instruction cache)
#pragma unroll 2
  for (int n = 0; n < TILE; ++n) {
    float2 nxt2;
    if ((n & 1) == 0) { w = rp[(n >> 1) + 1]; nxt2 = make_float2(w.x, w.y); }
    else nxt2 = make_float2(w.z, w.w);
    uint32_t word; float mu_e;
    bool em;
    if (SAMPLER == kRxSamplerHs) em = rx_sample_hs(p, r, cur, nxt, word, mu_e);
    else em = rx_sample<SAMPLER, SLICER>(p, r, cur, nxt, reinterpret_cast<const float2 *>(rp) + n, s_pe, word, mu_e);
    if (MODE != kWarm && em) {
      if (MODE == kOwned || MODE == kHead) {
        emit_word(e, e.n_out, word);
        ++e.n_out;
        if (MODE == kHead && e.hlog && e.n_head < kRxSeamLog) {
          e.hlog[e.n_head].t = t0 + (float)n + mu_e;
          e.hlog[e.n_head].sym = word >> 16;
          ++e.n_head;
        }
      } else {
        // Verification overlap: stored right after the owned symbols so that the
        // stitcher can extend this span by one symbol when needed.
        emit_word(e, e.n_out + e.n_tail, word);
        if (e.tlog && e.n_tail < kRxSeamLog) {
          e.tlog[e.n_tail].t = t0 + (float)n + mu_e;
          e.tlog[e.n_tail].sym = word >> 16;
        }
        ++e.n_tail;
      }
    }
    cur = nxt; nxt = nxt2;
  }
}

// End of a chunk: AGC / estimators / limits, measurement rows, end-of-span state.
template <int SAMPLER>
__device__ __forceinline__ void rx_chunk_close(const RxArgs &a, RxRun &r, uint64_t c, bool owned, uint32_t span,
                                               uint64_t own_end) {
  const RxParams &p = a.p;
  if (a.sampled && owned) {
    a.sampled_flag[c] = r.have_point ? 1u : 0u;
    if (r.have_point) a.sampled[c] = make_float2(r.s_re, r.s_im);
  }
  if (SAMPLER == kRxSamplerHs) rx_chunk_end_hs(p, r); else rx_chunk_end(p, r);
  // Measurements (sdr.h:904-913)
  r.meas_count += kRxChunk;
  while (r.meas_count >= p.meas_decimation) {
    r.meas_count -= p.meas_decimation;
    if (a.meas && owned) {
      const uint32_t k = atomicAdd(a.meas_count, 1u);
      if (k < a.max_meas) {
        float *m = a.meas + 4 * (size_t)k;
        m[0] = (float)c;
        m[1] = r.freq_tap;
        m[2] = __fsqrt_rn(r.est_insp);
        // est_sp / est_ep (or -1 when est_ep == 0): the host turns it into dB with glibc's
        // logf, like the reference (sdr.h:910-911), so that p_mer is bit-identical.
        m[3] = (r.est_ep != 0.0f) ? __fdiv_rn(r.est_sp, r.est_ep) : -1.0f;
      }
    }
  }
  if (c + 1 == own_end) store_state(a.state_end[span], r);
}

// Start state of a span that is warmed up: the carried loop state (frequency, AGC) with the timing / phase
// registers cleared; counters that are pure functions of the position are set for `run_begin`.
template <int SAMPLER>
__device__ __forceinline__ void rx_warm_state(const RxArgs &a, RxRun &r, uint64_t run_begin) {
  const RxParams &p = a.p;
  load_state(r, *a.warm_in);
  r.mu = 0.f; r.phase = 0.f;
  r.h0pr = r.h0pi = r.h0cr = r.h0ci = 0.f;
  r.h1pr = r.h1pi = r.h1cr = r.h1ci = 0.f;
  r.h2pr = r.h2pi = r.h2cr = r.h2ci = 0.f;
  // Position counters: state_in is valid at chunk a.state_chunk (0 unless a settling pass ran in front).
  const int64_t rel = (int64_t)run_begin - (int64_t)a.state_chunk;
  const int64_t dec = (int64_t)p.meas_decimation;
  int64_t mc = ((int64_t)a.state_in->meas_count + (rel % dec) * kRxChunk) % dec;
  if (mc < 0) mc += dec;
  r.meas_count = (uint32_t)mc;
  if (SAMPLER == 2) {
    // The tap-update throttle is a pure function of the position: first update at
    // chunk i0, then every ceil(n*16/128) chunks.
    const int R = p.rrc_n * 16, P = (R + kRxChunk - 1) / kRxChunk;
    const int p0 = a.state_in->rrc_update_phase;
    const int64_t i0 = (p0 <= kRxChunk) ? 0 : (p0 + kRxChunk - 1) / kRxChunk - 1;
    const int64_t rb = rel > 0 ? rel : 0;
    if (rb <= i0) r.rrc_update_phase = p0 - kRxChunk * (int)rb;
    else r.rrc_update_phase = R - kRxChunk * (int)((rb - i0 - 1) % P);
    r.rrc_f = __fdiv_rn(r.freqw, (float)p.rrc_sub);
  }
}

template <int SAMPLER, int SLICER, int TILE>
__device__ void rx_warp(const RxArgs &a, const uint32_t *span_list, uint32_t nlist, unsigned char *smem_rows,
                        uint32_t s_pe) {
  constexpr int kTile = TILE;
  constexpr int kTilesPerChunk = kRxChunk / kTile;
  constexpr int kRowBytes = RowCfg<SAMPLER, TILE>::kBytes;
  const RxParams &p = a.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * nwarps + warp;
  // Span of this lane.  Repair mode (span_list): lane g re-runs span span_list[g] exactly,
  // from the end state of its predecessor (a.state_end[span - 1]).
  const bool repair = (span_list != nullptr);
  const uint32_t g = warp_global * 32 + lane;
  uint32_t span;
  bool have_span;
  if (repair) { have_span = g < nlist; span = have_span ? span_list[g] : 0; }
  else { span = g; have_span = span < a.nspans; }
  const RxState *forced = nullptr;
  if (repair && have_span) forced = span ? a.state_end + (span - 1) : a.prev_end;

  const uint32_t S = a.span_chunks, W = a.warm_chunks;
  uint64_t own_begin = 0, own_end = 0, run_begin = 0, run_end = 0;
  RxRun r;
  load_state(r, *a.state_in);
  r.sg_re = r.sg_im = r.s_re = r.s_im = r.cp_re = r.cp_im = 0.f; r.have_point = 0;
  bool log_head_span = false;
  if (have_span) {
    own_begin = a.chunk0 + (uint64_t)span * S;
    own_end = own_begin + S;
    if (own_end > a.nchunks) own_end = a.nchunks;
    run_begin = (own_begin > W) ? own_begin - W : 0;
    // Verification overlap: as far as the data goes (avail_chunks == nchunks except in
    // time-sharded mode, where the last span also runs into the next rank's territory).
    run_end = own_end + kRxVerifyChunks;
    if (run_end > a.avail_chunks) run_end = a.avail_chunks;
    if (forced) {
      load_state(r, *forced);
      run_begin = own_begin;
    } else if (span == 0 && a.first_exact) {
      run_begin = own_begin;                 // the true state at chunk0 (batch start, or end of the settling pass)
    } else {
      rx_warm_state<SAMPLER>(a, r, run_begin);
    }
    log_head_span = a.head_log && (span > 0 || !a.first_exact);
  }
  // Common iteration space of the warp: local chunk index i, chunk c = base + i.
  // Lanes outside [run_begin, run_end) idle but keep the barrier protocol.
  int64_t base;          // chunk index of local iteration 0 for this lane
  uint32_t iters;
  if (repair) { base = (int64_t)run_begin; iters = S + kRxVerifyChunks; }
  else { base = (int64_t)(a.chunk0 + (uint64_t)span * S) - (int64_t)W; iters = W + S + kRxVerifyChunks; }
  // Local chunk indices (relative to base): [lb, le) is run, [ob, oe) is owned.
  const int lb = (int)((int64_t)run_begin - base), le = have_span ? (int)((int64_t)run_end - base) : lb;
  const int ob = (int)((int64_t)own_begin - base), oe = (int)((int64_t)own_end - base);

  // This lane's private row in each stage (shared-memory window of the warp).
  unsigned char *warp_rows = smem_rows + (size_t)warp * kStages * 32 * kRowBytes;
  const uint32_t total_tiles = iters * kTilesPerChunk;
  constexpr int kChunks16 = kRowBytes / 16;
  // Rows are fetched by the WARP: the 32 x kChunks16 16-byte pieces of a tile are dealt to the lanes
  // in row-major order, so one LDGSTS covers 32 / kChunks16 consecutive rows -- a few whole
  // 128-byte lines -- instead of 32 different lines (the per-lane version kept the L1 -> crossbar
  // request path 45 % busy, ncu round 1).  Row addresses travel by shuffle.
  const float2 *lane_x = a.x + base * (int64_t)kRxChunk;     // sample 0 of local chunk 0 (never dereferenced outside [lb, le))
  auto issue = [&](uint32_t tile) {
    const int st = (int)(tile % kStages);
    const int i = (int)(tile / kTilesPerChunk);
    const bool active = i >= lb && i < le;
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (act) {
      const uintptr_t src = reinterpret_cast<uintptr_t>(lane_x + (size_t)tile * kTile);
      const uint32_t slo = (uint32_t)src, shi = (uint32_t)(src >> 32);
      unsigned char *stage_base = warp_rows + (size_t)st * 32 * kRowBytes;
#pragma unroll
      for (int k = 0; k < kChunks16; ++k) {
        const int piece = k * 32 + lane, row = piece / kChunks16, q = piece % kChunks16;
        const uint32_t lo = __shfl_sync(0xffffffffu, slo, row), hi = __shfl_sync(0xffffffffu, shi, row);
        if ((act >> row) & 1u) {
          const unsigned char *rs = reinterpret_cast<const unsigned char *>(((uintptr_t)hi << 32) | lo);
          if (kRxCa) cp_async16_ca(stage_base + (size_t)row * kRowBytes + q * 16, rs + q * 16);
          else cp_async16(stage_base + (size_t)row * kRowBytes + q * 16, rs + q * 16);
        }
      }
    }
    cp_async_commit();
  };

  RxEmit e;
  e.out = a.sym_out + (size_t)span * a.span_cap;
  e.hlog = (log_head_span && have_span) ? a.head_log + (size_t)span * kRxSeamLog : nullptr;
  e.tlog = (a.tail_log && have_span) ? a.tail_log + (size_t)span * kRxSeamLog : nullptr;
  e.cap = a.span_cap;
  emit_init(e);

  for (int s = 0; s < kStages - 1; ++s) { if ((uint32_t)s < total_tiles) issue(s); else cp_async_commit(); }
  for (uint32_t tile = 0; tile < total_tiles; ++tile) {
    if (tile + kStages - 1 < total_tiles) issue(tile + kStages - 1); else cp_async_commit();
    cp_async_wait<kStages - 1>();
    __syncwarp();                           // rows were copied by other lanes
    const int st = (int)(tile % kStages);
    const int i = (int)(tile / kTilesPerChunk);
    const int tic = (int)(tile % kTilesPerChunk);
    if (i >= lb && i < le) {
      const int mode = (i < ob) ? kWarm : (i < oe ? ((e.hlog && i - ob < kRxVerifyChunks) ? kHead : kOwned) : kTail);
      if (tic == 0) {
        if (i == ob && a.state_begin) store_state(a.state_begin[span], r);   // state the span enters its own chunks with
        rx_chunk_begin(p, r, SAMPLER);
      }
      const float4 *rp = reinterpret_cast<const float4 *>(warp_rows + (size_t)st * 32 * kRowBytes + (size_t)lane * kRowBytes);
      switch (mode) {                        // warp-uniform except at the ends of the stream
        case kWarm: rx_tile<SAMPLER, SLICER, TILE, kWarm>(p, r, rp, s_pe, e, 0.f); break;
        case kOwned: rx_tile<SAMPLER, SLICER, TILE, kOwned>(p, r, rp, s_pe, e, 0.f); break;
        case kHead: rx_tile<SAMPLER, SLICER, TILE, kHead>(p, r, rp, s_pe, e, (float)((i - ob) * kRxChunk + tic * kTile)); break;
        default: rx_tile<SAMPLER, SLICER, TILE, kTail>(p, r, rp, s_pe, e, (float)((i - oe) * kRxChunk + tic * kTile)); break;
      }
      if (tic == kTilesPerChunk - 1)
        rx_chunk_close<SAMPLER>(a, r, (uint64_t)(base + i), mode == kOwned || mode == kHead, span, own_end);
    }
    __syncwarp();  // every lane is done with this stage before it is refilled
  }
  if (have_span) {
    emit_flush(e);
    RxSpanInfo inf;
    inf.n_out = e.n_out; inf.n_tail = e.n_tail; inf.n_head_logged = e.n_head; inf.pad = 0;
    a.info[span] = inf;
  }
}

// The 128 KB phase-error column into shared memory (SLICER 1), by the whole CTA.
__device__ __forceinline__ void load_pe16(int16_t *s_pe, const int16_t *g_pe) {
  const int4 *src = reinterpret_cast<const int4 *>(g_pe);
  int4 *dst = reinterpret_cast<int4 *>(s_pe);
  for (int i = threadIdx.x; i < kPe16Bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
}

template <int TILE, int SLICER>
__global__ void __launch_bounds__(SLICER ? 640 : 128)
k_rx(RxArgs a, const uint32_t *span_list, uint32_t nlist) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t s_pe = 0;                         // shared-space byte address of the phase-error column
  unsigned char *rows = smem;
  if (SLICER == 1) {
    load_pe16(reinterpret_cast<int16_t *>(smem), a.p.pe16);
    s_pe = smem_u32(smem);
    rows = smem + kPe16Bytes;
  }
  if (a.p.sampler == 0) rx_warp<0, SLICER, TILE>(a, span_list, nlist, rows, s_pe);
  else if (a.p.sampler == 1) rx_warp<1, SLICER, TILE>(a, span_list, nlist, rows, s_pe);
  else if (SLICER == 0 && a.p.sampler == kRxSamplerHs) rx_warp<kRxSamplerHs, 0, TILE>(a, span_list, nlist, rows, s_pe);
  else rx_warp<2, SLICER, TILE>(a, span_list, nlist, rows, s_pe);
}

// ------------------------------------------------------------------ the serial lane
// EXACT mode, the AGC settling pass of FAST mode and every other single-span run: ONE lane walks the chunks in
// the reference's order; the other 31 lanes of its warp only move data (whole chunks, 16 bytes per lane and
// copy, double buffered), so the recurrence never waits for memory and owns every issue slot of its scheduler.
constexpr int kSerialChunkBytes = (kRxChunk + 8) * 8;   // one chunk + look-ahead (<= 6 samples), 16-byte multiple

template <int SAMPLER, int SLICER>
__device__ void rx_serial(const RxArgs &a, unsigned char *bufs, uint32_t s_pe) {
  constexpr int TILE = 8;
  constexpr int kTilesPerChunk = kRxChunk / TILE;
  const RxParams &p = a.p;
  const int lane = threadIdx.x & 31;
  const uint64_t c0 = a.chunk0, c1 = a.nchunks;
  RxRun r;
  load_state(r, *a.state_in);
  r.sg_re = r.sg_im = r.s_re = r.s_im = r.cp_re = r.cp_im = 0.f; r.have_point = 0;
  if (!a.first_exact) rx_warm_state<SAMPLER>(a, r, c0);
  RxEmit e;
  e.out = a.sym_out; e.hlog = nullptr; e.tlog = nullptr;
  e.cap = a.span_cap;
  emit_init(e);
  constexpr int kPieces = (kRxChunk + (SAMPLER == 2 ? 6 : 2)) * 8 / 16;
  auto issue = [&](uint64_t c) {
    const unsigned char *src = reinterpret_cast<const unsigned char *>(a.x + c * kRxChunk);
    unsigned char *dst = bufs + (size_t)(c & 1) * kSerialChunkBytes;
    for (int q = lane; q < kPieces; q += 32) cp_async16(dst + q * 16, src + q * 16);
    cp_async_commit();
  };
  if (c0 < c1) issue(c0);
  for (uint64_t c = c0; c < c1; ++c) {
    if (c + 1 < c1) issue(c + 1); else cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    if (lane == 0) {
      if (c == c0 && a.state_begin) store_state(a.state_begin[0], r);
      rx_chunk_begin(p, r, SAMPLER);
      const unsigned char *row = bufs + (size_t)(c & 1) * kSerialChunkBytes;
#pragma unroll 1
      for (int tic = 0; tic < kTilesPerChunk; ++tic)
        rx_tile<SAMPLER, SLICER, TILE, kOwned>(p, r, reinterpret_cast<const float4 *>(row + tic * TILE * 8), s_pe, e, 0.f);
      rx_chunk_close<SAMPLER>(a, r, c, true, 0, c1);
    }
    __syncwarp();
  }
  if (lane == 0) {
    emit_flush(e);
    RxSpanInfo inf;
    inf.n_out = e.n_out; inf.n_tail = 0; inf.n_head_logged = 0; inf.pad = 0;
    a.info[0] = inf;
  }
}

template <int SLICER>
__global__ void __launch_bounds__(128)
k_rx_serial(RxArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t s_pe = 0;
  unsigned char *bufs = smem;
  if (SLICER == 1) {
    load_pe16(reinterpret_cast<int16_t *>(smem), a.p.pe16);
    s_pe = smem_u32(smem);
    bufs = smem + kPe16Bytes;
  }
  if (threadIdx.x >= 32) return;             // warps 1..3 only helped to load the table
  if (a.p.sampler == 0) rx_serial<0, SLICER>(a, bufs, s_pe);
  else if (a.p.sampler == 1) rx_serial<1, SLICER>(a, bufs, s_pe);
  else if (SLICER == 0 && a.p.sampler == kRxSamplerHs) rx_serial<kRxSamplerHs, 0>(a, bufs, s_pe);
  else rx_serial<2, SLICER>(a, bufs, s_pe);
}

// ---------------------------------------------------------------- seam stitching

// One warp per seam: lanes stride over the logged symbols.  sb / se: state of the later span entering its own
// chunks and end state of the earlier span (or null).
__device__ RxSeam stitch_seam(const RxStitchArgs &a, const RxSeamSym *tail, uint32_t nt, const RxSeamSym *head,
                              uint32_t nh, const RxState *sb, const RxState *se, int lane) {
  RxSeam s;
  s.ok = 0; s.rot = 0; s.extend_prev = 0; s.skip_next = 0; s.compared = 0; s.mismatches = 0; s.ok_loose = 0;
  s.dphase = 0.f; s.dfreqw = 0.f; s.dmu = 0.f;
  if (nt >= 8 && nh >= 8) {
    // Align on symbol time: tail[it0 + i] <-> head[ih0 + i].
    const float half = 0.5f * a.omega;
    int it0 = 0, ih0 = 0;
    const float d = head[0].t - tail[0].t;
    if (d > half) { it0 = 1; s.extend_prev = 1; }        // next span missed the first symbol
    else if (d < -half) { ih0 = 1; s.skip_next = 1; }    // next span repeats the previous span's last symbol
    const int n = (int)min(nt - it0, nh - ih0);
    int best_rot = 0, best_mis = 1 << 30;
    bool time_ok = true;
    for (int rot = 0; rot < a.nrot; ++rot) {
      const uint8_t *perm = a.rot_perm + rot * a.nsymbols;
      int mis = 0;
      for (int i = lane; i < n; i += 32) {
        const RxSeamSym hs = head[ih0 + i], ts = tail[it0 + i];
        if (perm[hs.sym] != ts.sym) ++mis;
        if (rot == 0 && fabsf(hs.t - ts.t) > 0.25f * a.omega) time_ok = false;
      }
      for (int o = 16; o; o >>= 1) mis += __shfl_xor_sync(0xffffffffu, mis, o);
      if (mis < best_mis) { best_mis = mis; best_rot = rot; }
      if (best_mis * 16 <= n) break;   // good enough: the other rotations disagree on ~3/4 of the symbols
    }
    time_ok = __all_sync(0xffffffffu, time_ok);
    s.rot = best_rot; s.compared = n; s.mismatches = best_mis;
    // Tolerant rule: isolated disagreements are noise-level decision flips between two converged
    // loops (either span may be the one that differs from the serial reference); an unconverged or
    // rotated span disagrees on half or more of the symbols.
    s.ok_loose = (time_ok && n >= 8 && best_mis * 16 <= n) ? 1 : 0;
    bool state_ok = true;
    if (sb && se) {
      // Loop states on both sides: the later span's phase is in its own frame (rot * 65536/nrot away).
      const float sector = 65536.0f / (float)a.nrot;
      float dp = fmodf(sb->phase - se->phase - (float)best_rot * sector, sector);
      if (dp > 0.5f * sector) dp -= sector;
      if (dp < -0.5f * sector) dp += sector;
      s.dphase = dp;
      s.dfreqw = sb->freqw - se->freqw;
      s.dmu = sb->mu - se->mu - (float)(s.extend_prev - s.skip_next) * a.omega;
      if (a.tol_phase > 0.f && fabsf(dp) > a.tol_phase) state_ok = false;
      if (a.tol_freqw > 0.f && fabsf(s.dfreqw) > a.tol_freqw) state_ok = false;
    }
    s.ok = a.strict ? (s.ok_loose && best_mis == 0 && state_ok) : s.ok_loose;
  }
  return s;
}

__global__ void __launch_bounds__(128)
k_rx_stitch(RxStitchArgs a, const uint32_t *seam_list, uint32_t nlist) {
  const int lane = threadIdx.x & 31;
  uint32_t j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seam_list) { if (j >= nlist) return; j = seam_list[j]; }
  if (j + 1 >= a.nspans) return;
  const bool st = a.state_begin && a.state_end;
  const RxSeam s = stitch_seam(a, a.tail_log + (size_t)j * kRxSeamLog, min(a.info[j].n_tail, (uint32_t)kRxSeamLog),
                               a.head_log + (size_t)(j + 1) * kRxSeamLog,
                               min(a.info[j + 1].n_head_logged, (uint32_t)kRxSeamLog),
                               st ? a.state_begin + (j + 1) : nullptr, st ? a.state_end + j : nullptr, lane);
  if (lane == 0) a.seams[j] = s;
}

__global__ void k_rx_stitch_pair(RxStitchArgs a, const RxSeamSym *tail, uint32_t n_tail, const RxState *prev_end, RxSeam *out) {
  const RxSeam s = stitch_seam(a, tail, min(n_tail, (uint32_t)kRxSeamLog), a.head_log,
                               min(a.info[0].n_head_logged, (uint32_t)kRxSeamLog),
                               (a.state_begin && prev_end) ? a.state_begin : nullptr, prev_end, threadIdx.x & 31);
  if (threadIdx.x == 0) *out = s;
}

__global__ void k_rx_compact(RxCompactArgs a, uint64_t total) {
  // One block row per span; threads copy with the span's rotation applied.
  const uint32_t span = blockIdx.x;
  const uint64_t base = a.span_offset[span];
  const uint64_t n = a.span_offset[span + 1] - base;
  const uint32_t *src = a.sym_in + (size_t)span * a.span_cap + a.span_skip[span];
  const uint8_t *perm = a.rot_perm + (size_t)a.span_rot[span] * a.nsymbols;
  const bool identity = (a.span_rot[span] == 0);
  auto fix = [&](uint32_t w) {
    if (identity) return w;
    const uint32_t sym = (w >> 16) & 0xffu;
    return (w & 0xffffu) | ((uint32_t)perm[sym] << 16);
  };
  uint32_t *dst = a.sym_out + base;
  // Head up to the first 16-byte boundary of the destination, then 16-byte stores (the source is only 4-byte
  // aligned relative to it: four 4-byte loads per thread, still contiguous across the warp), then the tail.
  const uint64_t head = min((uint64_t)((16u - ((uint32_t)reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) / 4u, n);
  for (uint64_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = fix(src[i]);
  const uint64_t nvec = (n - head) / 4;
  for (uint64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint64_t i = head + 4 * v;
    const uint4 w = make_uint4(fix(__ldcs(src + i)), fix(__ldcs(src + i + 1)), fix(__ldcs(src + i + 2)), fix(__ldcs(src + i + 3)));
    *reinterpret_cast<uint4 *>(dst + i) = w;
  }
  for (uint64_t i = head + 4 * nvec + threadIdx.x; i < n; i += blockDim.x) dst[i] = fix(src[i]);
}

#include "k_ctl_rx.cuh"   // k_rx_plan_local, k_rx_plan_apply


// Mean |x|^2 of the first n samples.
__global__ void __launch_bounds__(256)
k_rx_power(const float2 *x, uint32_t n, float *out) {
  __shared__ float part[8];
  float acc = 0.f;
  for (uint32_t i = threadIdx.x; i < n; i += 256) { const float2 v = x[i]; acc += v.x * v.x + v.y * v.y; }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    *out = n ? t / (float)n : 0.f;
  }
}

}  // namespace

cudaError_t launch_rx_power(const float2 *x, uint32_t n, float *out, cudaStream_t st) {
  k_rx_power<<<1, 256, 0, st>>>(x, n, out);
  return cudaGetLastError();
}

namespace {
constexpr int kTile = 8;   // samples per staged tile (16 was measured 4 % slower in round 1)

int env_int(const char *name, int dflt, int lo, int hi) {
  const char *e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return (v < lo || v > hi) ? dflt : v;
}

size_t rx_row_bytes(int sampler) { return sampler == 2 ? RowCfg<2, kTile>::kBytes : RowCfg<1, kTile>::kBytes; }
}  // namespace

// Warps per CTA of the span kernel.  Slicer 1 (QPSK, phase-error column in shared memory): one fat CTA per SM,
// LDVB_RX_WARPS warps (default 12: 128 KB of table + 60 KB of rows; measured 1.60 ms against 2.35 ms with 16 warps, whose
// 208 KB of shared memory leave the trig16 gathers only 28 KB of L1); slicer 0: 4 warps, three CTAs per SM.
int rx_warps_per_cta(int slicer) {
  static const int w1 = env_int("LDVB_RX_WARPS", 12, 1, 20);
  return slicer == 1 ? w1 : 4;
}

// Lanes (spans) of one full wave of the span kernel on the current device.
uint64_t rx_resident_lanes(int slicer) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (uint64_t)sms * (slicer == 1 ? rx_warps_per_cta(1) : 12) * 32;
}

cudaError_t launch_rx(const RxArgs &a, const uint32_t *span_list, uint32_t nlist, cudaStream_t st) {
  if (a.nspans == 0 || (span_list && !nlist)) return cudaSuccess;
  const int slicer = a.p.slicer;
  if (slicer == 1 && !a.p.pe16) return cudaErrorInvalidValue;
  // One span, no warm-up: the serial lane (EXACT mode, settling pass).
  if (!span_list && a.nspans == 1 && a.warm_chunks == 0 && !a.head_log && !a.tail_log) {
    const size_t smem = (slicer == 1 ? (size_t)kPe16Bytes : 0) + 2 * (size_t)kSerialChunkBytes;
    static PerDeviceMark configured;
    if (slicer == 1 && configured.need(1)) {
      cudaError_t e = cudaFuncSetAttribute(k_rx_serial<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured.commit(1);
    }
    if (slicer == 1) k_rx_serial<1><<<1, 128, smem, st>>>(a);
    else k_rx_serial<0><<<1, 32, smem, st>>>(a);
    return cudaGetLastError();
  }
  const int warps = rx_warps_per_cta(slicer);
  const unsigned per_block = (unsigned)warps * 32;
  const unsigned lanes = span_list ? nlist : a.nspans;
  const size_t rows = (size_t)warps * kStages * 32 * rx_row_bytes(a.p.sampler);
  if (slicer == 1) {
    const size_t smem = (size_t)kPe16Bytes + rows;
    static PerDeviceMark configured;
    if (configured.need(smem)) {
      cudaError_t e = cudaFuncSetAttribute(k_rx<kTile, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured.commit(smem);
    }
    k_rx<kTile, 1><<<(lanes + per_block - 1) / per_block, per_block, smem, st>>>(a, span_list, nlist);
    return cudaGetLastError();
  }
  static PerDeviceMark configured0;
  if (configured0.need(1)) {
    // Leave most of the SM's unified cache to L1: the two 512 KB tables are read through it,
    // and their hot lines (current carrier phase, constellation clusters) must stay resident.
    // Measured (round 1): 8-sample tiles (20 KB of rows per CTA, 3 CTAs per SM) with a 30-40 %
    // shared-memory carve-out; every split that gives L1 less than half is 30-130 % slower.
    static const int carveout = env_int("LDVB_RX_CARVEOUT", 35, 0, 100);
    cudaError_t e = cudaFuncSetAttribute(k_rx<kTile, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
    if (e != cudaSuccess) return e;
    configured0.commit(1);
  }
  k_rx<kTile, 0><<<(lanes + per_block - 1) / per_block, per_block, rows, st>>>(a, span_list, nlist);
  return cudaGetLastError();
}

cudaError_t launch_rx_stitch(const RxStitchArgs &a, const uint32_t *seam_list, uint32_t nlist, cudaStream_t st) {
  if (a.nspans < 2 || (seam_list && !nlist)) return cudaSuccess;
  const unsigned n = seam_list ? nlist : a.nspans - 1;
  k_rx_stitch<<<(n + 3) / 4, 128, 0, st>>>(a, seam_list, nlist);
  return cudaGetLastError();
}

cudaError_t launch_rx_stitch_pair(const RxStitchArgs &a, const RxSeamSym *tail, uint32_t n_tail, const RxState *prev_end,
                                  RxSeam *out, cudaStream_t st) {
  k_rx_stitch_pair<<<1, 32, 0, st>>>(a, tail, n_tail, prev_end, out);
  return cudaGetLastError();
}

cudaError_t launch_rx_plan(const RxSpanInfo *info, const RxSeam *seams, uint32_t nspans, uint32_t span_cap, int nrot,
                           int rot0, uint32_t skip0, uint64_t *span_offset, uint32_t *span_skip, uint8_t *span_rot,
                           uint64_t *result, cudaStream_t st) {
  // (the CTA totals live behind the nspans + 1 offsets: the handle allocates span_offset with room for them)
  const uint32_t nblk = std::max(1u, (nspans + 1023u) / 1024u);
  unsigned long long *totals = reinterpret_cast<unsigned long long *>(span_offset + (size_t)nspans + 1);
  unsigned long long *res = reinterpret_cast<unsigned long long *>(result);
  cudaError_t e = cudaMemsetAsync(result, 0, 9 * sizeof(uint64_t), st);
  if (e != cudaSuccess) return e;
  k_rx_plan_local<<<nblk, 1024, 0, st>>>(info, seams, nspans, span_cap, nrot, rot0, skip0, span_offset, span_skip, span_rot, totals, res);
  k_rx_plan_apply<<<nblk, 1024, 0, st>>>(nspans, nrot, span_offset, span_rot, totals, res);
  return cudaGetLastError();
}

cudaError_t launch_rx_compact(const RxCompactArgs &a, uint64_t total, cudaStream_t st) {
  if (a.nspans == 0 || total == 0) return cudaSuccess;
  // Spans are short (a few thousand symbols): one 256-thread block per span.
  k_rx_compact<<<a.nspans, 256, 0, st>>>(a, total);
  return cudaGetLastError();
}

}  // namespace ldvb
