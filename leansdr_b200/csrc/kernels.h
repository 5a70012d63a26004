// kernels.h -- launch interfaces of the sm_100a kernels (host side).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace ldvb {

// Two-part view of an input stream: logical element i lives in `head` when
// i < c0 or main == nullptr, else in main[i - c0].  `head` holds the samples
// carried from the previous batch followed by the first samples of the new one,
// so that a device-resident batch (`main`) is processed in place, without a copy.
struct RawSrc {
  const void *head; uint64_t head_count;
  const void *main; uint64_t c0;
};

// ------------------------------------------------------------------ K1 frontend
struct FrontendArgs {
  RawSrc src;             // input stream (both parts 16-byte aligned)
  int fmt;                // 0 u8, 1 s8, 2 u16, 3 s16, 4 f32 (scaled), 5 cf32 (as is)
  float scale;            // --float-scale (fmt 4)
  const float *rot_lut;   // 65536 cos then 65536 sin, or nullptr
  uint32_t rot_index0;    // rotator index of sample 0
  const float2 *taps;     // ntaps shifted complex taps, ntaps == 0: copy/decimate only
  uint32_t ntaps, decim;
  int real_taps;          // every tap has a zero imaginary part (no retune in force)
  float2 *out;
  uint64_t count;         // outputs to produce
  // filled by launch_frontend
  uint32_t tile_out, bytes_per_sample, max_raw_bytes;
};
int frontend_bytes_per_sample(int fmt);
cudaError_t launch_frontend(FrontendArgs a, cudaStream_t st);

// --------------------------------------------------------------------- K2 notch
constexpr int kNotchN = 4096;
constexpr int kNotchMaxSlots = 4;

struct NotchSlotState { int32_t bin; float est_re, est_im; int32_t pad; };
struct NotchState {
  int32_t phase;          // samples since the last detect() (sdr.h:66-70)
  float gain;
  NotchSlotState slot[kNotchMaxSlots];
};

// Detection: one CTA per detect point runs the 4096-point radix-2 inverse FFT
// and the peak search of auto_notch::detect() (sdr.h:76-118).
struct NotchDetectArgs {
  RawSrc src; int fmt; float scale;           // same sample access as the front end
  const uint64_t *block_index;                // [ndetect] 4096-block index of each detect point
  int ndetect, nslots;
  const float2 *twiddle_rev;                  // [4096] omega_rev (dsp.h:70-76)
  int32_t *bins_out;                          // [ndetect][nslots]
};
cudaError_t launch_notch_detect(NotchDetectArgs a, cudaStream_t st);

// Builds expj tables for an epoch: [nslots][4096] (sdr.h:104-108); computed on
// the host (glibc cosf/sinf) and uploaded, the kernel only applies them.
struct NotchEpoch {
  uint64_t first_block;   // first 4096-block of the batch that uses this epoch's tables
  int32_t bin[kNotchMaxSlots];
  int32_t reset[kNotchMaxSlots];  // slot estimate is zeroed at first_block (bin changed)
  uint32_t table_index[kNotchMaxSlots];  // which uploaded expj table each slot uses
};

struct NotchApplyArgs {
  RawSrc src; int fmt; float scale;
  float2 *out;
  uint64_t nblocks;                // 4096-sample blocks to process
  int nslots;
  float k, gain;
  float w_block;                   // (1-k)^4096, for the start-state guess
  const float2 *expj_tables;       // [ntables][4096]
  const NotchEpoch *epochs;        // device, sorted by first_block
  int nepochs;
  uint64_t block0;                 // first owned block (time-sharded mode: blocks before it are halo)
  int first_exact;                 // segment 0 starts from the exact carried state (state_in)
  uint32_t seg_blocks;             // blocks owned per segment
  uint32_t warm_blocks;            // warm-up blocks before a speculative segment
  uint32_t nsegs;
  const NotchState *state_in;      // exact state at block 0
  float2 *seg_entry;               // [nsegs][slots] state at segment start (after warm-up)
  float2 *seg_exit;                // [nsegs][slots] state at segment end
  uint8_t *seg_exact;              // [nsegs] 1 when the segment started from a known-exact state
};
// guess: [nblocks][kNotchMaxSlots] per-block weighted sums written by launch_notch_guess;
// launch_notch_apply assembles the start states of speculative segments from them.
cudaError_t launch_notch_guess(const NotchApplyArgs &a, float2 *guess, const float *weights, cudaStream_t st);
// Counts the segments whose entry state differs from their predecessor's exit state.
cudaError_t launch_notch_verify(const NotchApplyArgs &a, uint32_t *nfail, cudaStream_t st);
// seg_list == nullptr: every segment (warm-up from `guess`).  Otherwise the listed segments
// are re-run exactly from the exit state of their predecessors (a.seg_exit).
cudaError_t launch_notch_apply(NotchApplyArgs a, const uint32_t *seg_list, uint32_t nlist,
                               const float2 *guess, cudaStream_t st);

// Fused auto_notch + fir_filter (k_notchfir.cu): the notch chain of a CTA's 32 segments is walked by one warp,
// the data-parallel work around it by eight others, and -- when the low-pass that follows has decimation 1 and at
// most kFirFuseMaxTaps taps -- the FIR is applied on the store path, so that the notched stream never reaches HBM.
constexpr int kFirFuseMaxTaps = 32;
constexpr int kNotchFirRows = 16;               // segments per CTA of k_notch_fir
constexpr int kNotchEdge = 2 * kFirFuseMaxTaps;   // float2 per segment: [first N-1 | (at kFirFuseMaxTaps) last N] notched samples
struct NotchFirArgs {
  NotchApplyArgs n;             // n.out is used only when fir_n == 0 (plain notch)
  int fir_n;                    // 0: no FIR
  int real_taps;                // every tap has a zero imaginary part (no retune in force)
  const float2 *taps;           // [fir_n] shifted taps (dsp.h:270-280)
  float2 *y;                    // y[k], k = 0 .. carry + nblocks*4096 - fir_n - 1 (dsp.h:250-256)
  uint32_t carry;               // notched samples carried in front of the batch: 0 (first batch) or fir_n
  const float2 *carry_in;       // [carry]
  float2 *carry_out;            // [fir_n] the last fir_n notched samples of this batch (may alias carry_in)
  float2 *edge;                 // [nsegs][kNotchEdge]
  const uint64_t *dump_blocks;  // [ndump] sorted: blocks whose notched samples cnr_fft / spectrum will read
  int ndump;
  float2 *dump;                 // [ndump][4096]
};
cudaError_t launch_notch_fir(const NotchFirArgs &a, const uint32_t *seg_list, uint32_t nlist, const float2 *guess, cudaStream_t st);
// The fir_n - 1 outputs across every segment boundary (and the batch start), then the new carry.
cudaError_t launch_fir_edges(const NotchFirArgs &a, cudaStream_t st);

// ------------------------------------------------------- K8 cnr_fft / spectrum
// [carry | rest]: `carry` holds cf32 samples (converted, not rotated) kept from the previous
// batch, `rest` is the batch in its input format starting at element rest_off.
struct MeasSrc {
  const float2 *carry; uint64_t carry_count;
  RawSrc rest; uint64_t rest_off; int fmt; float scale;
  const float *rot_lut; uint32_t rot_index0;   // rotator index of element 0 (or no rotator)
};
struct MeasArgs {
  MeasSrc src;
  const uint64_t *point_start;   // [npoints] element index of each measured block
  int npoints, logn;             // block = 2^logn samples (12: cnr_fft, 10: spectrum)
  const float2 *twiddle_rev;     // [4096] omega_rev of the 4096-point engine
  float *power;                  // [npoints][n]
};
cudaError_t launch_meas_power(const MeasArgs &a, cudaStream_t st);
struct MeasEmaArgs {
  const float *power; int npoints, n;
  float kavg;
  float *avg; int *have;         // carried average (device)
  int bwslots, icf;              // cnr_fft band geometry (bwslots == 0: none)
  float *sums;                   // [npoints][3] c2+n2, noise left, noise right
  float *rows;                   // [npoints][n] average after each measurement, or null
};
cudaError_t launch_meas_ema(const MeasEmaArgs &a, cudaStream_t st);
cudaError_t launch_meas_save(const MeasSrc &src, uint64_t start, uint32_t count, float2 *dst, cudaStream_t st);

// ------------------------------------------------------------------ K3 receiver
struct CstlnCellDev { int16_t cost, symbol, phase_error, pad; };

constexpr int kPeFoldPitch = 129;   // |Q| = 0..128
struct RxParams {
  const CstlnCellDev *cstln;   // [65536]
  const float2 *trig;          // [65536]
  int8_t sym_re[256], sym_im[256];
  int nsymbols, sampler;
  float omega, min_freqw, max_freqw;
  float freq_alpha, freq_beta, gain_mu, kest;
  int allow_drift;
  uint32_t meas_decimation;
  // rrc sampler
  const float *rrc_coeffs; int rrc_n, rrc_sub;
  // --hs: fast_qpsk_receiver (sampler == kRxSamplerHs).  The integer loop state travels in the
  // float fields of RxState (phase: u16, freqw and the limits: integers below 2^24, hist: u8) --
  // every value is exactly representable, so carry / warm-up / seam plumbing is shared.
  const uint32_t *hs_polar; const uint16_t *hs_rect; const uint16_t *hs_sincos;
  const int16_t *pe16;       // slicer 1: phase_error column folded over Q, [I & 0xff][|Q|] with pitch kPeFoldPitch
                             // (pe(I, -Q) == -pe(I, Q), verified at create time); copied to shared memory
  int slicer;                // 0: cell table gather (any constellation); 1: QPSK arithmetic + pe16 (k_rx.cu)
  long long hs_freq_beta;    // (signed long)(0.0012*256*65536/omega*pll_adjustment), sdr.h:1002
};
constexpr int kRxSamplerHs = 3;

// dvb_deconvol_sync_hard (dvb.h:612-707): per 64-byte chunk; see k_fec.cu.
struct HsDeconvArgs {
  const uint32_t *symbols;   // softsymbol words, hard symbol in bits 16..17
  uint64_t nchunks;          // chunks of 512 symbols / 64 bytes
  uint64_t hist;             // the 32 symbols in front of symbols[0], 2 bits each, newest in the LSBs
  int hist_valid;            // how many of them exist (0 at the start of a stream)
  int resync_phase, resync_period, locked;
  uint32_t *errors;          // [ngroups][4] scratch
  uint8_t *lock_of_chunk;    // [ngroups] scratch: alignment in force after the vote of group g
  uint8_t *out;              // nchunks * 64 bytes
  int32_t *state_out;        // {locked after the batch}
};
cudaError_t launch_hs_deconv(const HsDeconvArgs &a, cudaStream_t st, int *launches);

// 22 words, same layout as ldvb_get_rx_state (include/leandvb_b200.h).
struct RxState {
  float mu, phase, freqw, est_insp, agc_gain, est_sp, est_ep;
  float hist[12];          // hist[k] = {p.re, p.im, c.re, c.im}
  float samp_freqw, freq_tap;
  uint32_t meas_count;
  int32_t rrc_update_phase;  // fir_sampler::update_freq_phase (sdr.h:688)
  float rrc_f;               // freqw/subsampling at the last tap update (sdr.h:679)
};

struct RxSeamSym { float t; uint32_t sym; };  // time relative to the seam, hard symbol

struct RxSpanInfo {
  uint32_t n_out;          // symbols emitted in the owned region (sample index < end)
  uint32_t n_tail;         // symbols emitted in the verification overlap after the end
  uint32_t n_head_logged;  // entries in the head log
  uint32_t pad;
};

constexpr int kRxChunk = 128;
constexpr int kRxSeamLog = 192;        // log entries per seam side
constexpr int kRxVerifyChunks = 1;     // chunks of overlap after a span end

struct RxArgs {
  RxParams p;
  const float2 *x;           // preprocessed stream, chunk c starts at x[c*128]
  uint64_t nchunks;          // end of the owned chunks (spans cover [chunk0, nchunks))
  uint64_t avail_chunks;     // chunks present in x (>= nchunks; == nchunks outside time-sharded mode)
  uint64_t chunk0;           // first owned chunk (time-sharded mode: chunks before it are halo)
  int first_exact;           // span 0 starts at chunk0 from the exact carried state (state_in), without warm-up
  const RxState *prev_end;   // repair mode: end state of the span before span 0 (previous rank), or null
  uint32_t span_chunks;      // owned chunks per span (exact mode: >= nchunks)
  uint32_t warm_chunks;      // warm-up chunks (0 in exact mode)
  uint32_t nspans;
  uint32_t span_cap;         // symbol capacity of one span's output region
  const RxState *state_in;   // exact/carried state at chunk `state_chunk` (span 0 starts from it when first_exact)
  uint64_t state_chunk;      // chunk at which state_in's position counters (meas_count, rrc phase) are valid
  const RxState *warm_in;    // loop state (freqw, AGC) that warm-ups start from; usually == state_in
  uint32_t *sym_out;         // [nspans][span_cap] softsymbols {cost:16, symbol:8, 0}
  RxSpanInfo *info;          // [nspans]
  RxState *state_end;        // [nspans] state at the span's nominal end
  RxState *state_begin;      // optional [nspans]: state with which the span enters its own chunks (after the warm-up)
  RxSeamSym *head_log;       // [nspans][kRxSeamLog]
  RxSeamSym *tail_log;       // [nspans][kRxSeamLog]
  float2 *sampled;           // optional [nchunks] tap (exact mode), may be null
  uint32_t *sampled_flag;    // optional [nchunks]
  float *meas;               // optional [max_meas][4] {chunk, freq_tap, ss, mer}
  uint32_t *meas_count;      // optional counter
  uint32_t max_meas;
};
// span_list == nullptr: all spans.  Otherwise the listed spans are re-run exactly from the end
// state of their predecessors (a.state_end[span - 1]) -- seam repair.
cudaError_t launch_rx(const RxArgs &a, const uint32_t *span_list, uint32_t nlist, cudaStream_t st);
int rx_warps_per_cta(int slicer);
uint64_t rx_resident_lanes(int slicer);   // lanes (spans) of one full wave of the span kernel on the current device

struct RxSeam {
  int32_t ok;              // verification passed (rule in force: strict or tolerant, see RxStitchArgs)
  int32_t rot;             // rotation of span j+1 relative to span j (units of 360/nrot)
  int32_t extend_prev;     // span j keeps this many tail symbols (0/1)
  int32_t skip_next;       // span j+1 drops this many head symbols (0/1)
  int32_t compared;        // symbols compared
  int32_t mismatches;      // hard decisions that differ in the overlap (after de-rotation)
  int32_t ok_loose;        // the tolerant rule alone: aligned in time and <= 1/16 mismatches
  float dphase, dfreqw, dmu;  // state of span j+1 entering its chunks minus end state of span j (phase modulo the
                              // rotational ambiguity, phase units / freqw units / samples); 0 when not recorded
};
struct RxStitchArgs {
  const RxSpanInfo *info; const RxSeamSym *head_log; const RxSeamSym *tail_log;
  uint32_t nspans; int nrot; int nsymbols;
  const uint8_t *rot_perm;   // [nrot][nsymbols] device
  float omega;
  RxSeam *seams;             // [nspans-1]
  // strict: a seam verifies only when EVERY compared hard decision agrees (and the loop states agree within the
  // tolerances below); otherwise the tolerant rule (<= 1/16 mismatches) decides.
  int strict;
  const RxState *state_begin, *state_end;   // [nspans] or null: loop states on both sides of each seam
  float tol_phase, tol_freqw;               // |dphase|, |dfreqw| bounds for `ok` (0: not checked)
};
cudaError_t launch_rx_stitch(const RxStitchArgs &a, const uint32_t *seam_list, uint32_t nlist, cudaStream_t st);
// One seam between an imported tail log (previous rank) and span 0's head log.
cudaError_t launch_rx_stitch_pair(const RxStitchArgs &a, const RxSeamSym *tail, uint32_t n_tail, const RxState *prev_end,
                                  RxSeam *out, cudaStream_t st);

// Scan over the seams: offsets / skips / cumulative rotations of every span (two grids of 1024-span CTAs,
// k_ctl_rx.cuh).  span_offset holds nspans + 1 offsets AND, behind them, 2 * ceil(nspans / 1024) words of CTA totals:
// allocate 8 * (nspans + 1) + 16 * (nspans / 1024 + 2) bytes.  result: [9] words, zeroed by the launcher.
// result[0] = seams that failed verification, [1] = symbols kept, [2] = rotation of the
// last span, [3] = spans that overflowed their capacity, [4] = verified seams with at least one
// mismatching hard decision (tolerant rule), [5..7] = max |dphase|, |dfreqw|, |dmu| over the seams (float bits),
// [8] = seams that fail the tolerant rule too (spans that did not converge).
cudaError_t launch_rx_plan(const RxSpanInfo *info, const RxSeam *seams, uint32_t nspans, uint32_t span_cap, int nrot,
                           int rot0, uint32_t skip0, uint64_t *span_offset, uint32_t *span_skip, uint8_t *span_rot,
                           uint64_t *result, cudaStream_t st);

// Concatenates the span outputs into one contiguous softsymbol stream, applying
// each span's cumulative rotation to the hard symbol.
struct RxCompactArgs {
  const uint32_t *sym_in; uint32_t span_cap; uint32_t nspans;
  const uint64_t *span_offset;   // [nspans+1] exclusive prefix of kept counts (device)
  const uint32_t *span_skip;     // [nspans]
  const uint8_t *span_rot;       // [nspans] cumulative rotation
  const uint8_t *rot_perm; int nsymbols;
  uint32_t *sym_out;             // appended after the carried symbols
};
cudaError_t launch_rx_compact(const RxCompactArgs &a, uint64_t total, cudaStream_t st);

// Mean |x|^2 of the first n samples (one CTA): the level check in front of the FAST spans.
cudaError_t launch_rx_power(const float2 *x, uint32_t n, float *out, cudaStream_t st);

// -------------------------------------------------------------- K4 deconvolution
struct DeconvArgs {
  const uint32_t *symbols;   // softsymbols, element 0 = first unread symbol
  uint64_t nbytes;           // bytes to produce
  uint64_t reg_in; int n_in; // carried shift register of the locked hypothesis (dvb.h:297-303)
  uint64_t out_acc; int n_out;
  uint8_t hyp[4];            // symbol&3 -> IQ bits of the locked hypothesis
  int punctperiod, punctweight;
  uint64_t deconv[8];
  uint8_t *out;
  // --fastlock (dvb.h:391-412, 428-442): when err_out is set the kernel writes no bytes; it counts,
  // over every bit group that STARTS inside the window, the bits on which the alternate polynomials
  // deconv2 disagree with deconv (readerrors on the auxiliary register reg_in/n_in/n_out).
  uint64_t deconv2[8];
  unsigned long long *err_out;
};
// carry_out (device, 5 x uint64): register, n_in, accumulator, n_out, symbols consumed.
cudaError_t launch_deconv_carry(const DeconvArgs &a, uint64_t nsym, uint64_t *carry_out, cudaStream_t st);

// --------------------------------------------------------------------- K5 Viterbi
struct VitDecState { int32_t cost[64]; uint64_t path[64]; int32_t bank, pad; };
struct VitCtl { int32_t current_sync, resync_phase; };
struct VitArgs {
  const uint32_t *symbols;   // softsymbols, element 0 = first unread symbol
  uint64_t nchunks;          // chunks of 128 FEC blocks (dvb.h:1372-1373)
  int bits_in, bits_out, bps, nshifts, nsyncs, ncs, nsymbols;
  int path_nbits, path_depth, path32, resync_period;
  const uint8_t *trellis_pred, *trellis_us;   // [64][ncs], pred == 65: no branch
  const uint8_t *maps;       // [nsyncs][nsymbols] (dvb.h:1336-1351)
  const int32_t *shifts;     // [nsyncs]
  VitDecState *state;        // [nsyncs]
  VitCtl *ctl;
  uint8_t *out;              // nchunks * 128 * bits_in / 8 bytes
};
// Time segments of the Viterbi stage (see k_viterbi.cu): segment g covers chunks
// [seg_start[g], seg_start[g+1]); every boundary but the first is a re-sync chunk.
struct VitSegArgs {
  const uint64_t *seg_start;   // [nseg + 1] (device)
  uint32_t nseg;
  const uint32_t *list;        // null: all segments (g > 0 start cold, warmed up); else the segments to re-run
  uint32_t nlist;              //       exactly from exit[g-1]
  uint32_t warm_chunks;        // warm-up of the current decoder (of all decoders when resync_period == 1)
  uint32_t warm_others;        // resync_period > 1: re-sync chunks every decoder warms up on (0 with warm_chunks == 0:
                               // no warm-up at all, a test knob: every cold segment fails verification)
  int phase0;                  // resync_phase at chunk 0
  int nb;                      // rescan entries per state = vit_rescan_entries(bits_in)
  int full;                    // every state is a predecessor of every state (vit_trellis_is_full): k_viterbi<kVitFull>
  VitDecState *entry, *exit;   // [nseg][nsyncs]
  VitCtl *ctl_entry, *ctl_exit;// [nseg]
};
int vit_rescan_entries(int bits_in);
bool vit_trellis_is_full(const uint8_t *pred /* [64][ncs] host */, int ncs, int bits_in);
int vit_resident_segments(int ncs, int bits_in, int nsyncs);   // CTAs of one full wave on the current device (0: unknown)
// nblocks = nseg (list == null) or nlist.
cudaError_t launch_viterbi(const VitArgs &a, const VitSegArgs &sg, uint32_t nblocks, cudaStream_t st);
cudaError_t launch_vit_verify(const VitSegArgs &sg, int nsyncs, uint8_t *ok, uint32_t *nfail, cudaStream_t st);
cudaError_t launch_vit_commit(const VitArgs &a, const VitSegArgs &sg, cudaStream_t st);

// ---------------------------------------------------------------- K6/K7 framing
struct SyncState {
  int32_t synchronized, bitphase, polarity, phase8;
  int32_t next_sync_count;
  uint32_t lock_timeleft;
  uint64_t locktime;
  int32_t report_state;
  int32_t fastlock;          // run_searching_fast instead of run_searching (dvb.h:751, 781-796)
  int32_t resync_period;     // dvb.h:717: search every resync_period-th packet position
  int32_t resync_phase;
};

// Outcome of one pass of the MPEG sync tracker over the byte stream.
struct SyncResult {
  SyncState st;              // state after the pass
  uint64_t consumed;         // bytes consumed from the input
  uint64_t produced;         // aligned bytes to append to mpegbytes (multiple of 204)
  int32_t need_next_sync;    // third fruitless sweep completed: switch hypothesis (dvb.h:771-778)
  int32_t events;            // lock transitions in this pass
  int32_t event_val[16];
  uint64_t event_pos[16];
};

cudaError_t launch_sync_flags(const uint8_t *bytes, uint64_t npackets, const SyncState *st_dev,
                              uint32_t *bad_words, cudaStream_t st);
cudaError_t launch_sync_track(const uint8_t *bytes, uint64_t nbytes, const SyncState *st_dev,
                              const uint32_t *bad_words, uint64_t npackets_flagged, SyncResult *res,
                              cudaStream_t st);
cudaError_t launch_realign(const uint8_t *bytes, uint64_t n, int bitphase, int polarity, uint8_t *out,
                           cudaStream_t st);

// Deinterleaver + RS(204,188) + flags; warp per packet.
struct DeintRsArgs {
  const uint8_t *mpeg;       // aligned byte stream, element 0 = oldest byte kept (history)
  uint64_t npackets;         // packets to decode: packet p reads mpeg[204*p .. 204*p+2447]
  const uint8_t *gf_exp, *gf_log;   // device tables (512, 256)
  uint8_t *rs_out;           // optional tap: deinterleaved 204-byte packets (may be null)
  uint8_t *rts_out;          // 188-byte packets
  int32_t *flags;            // [npackets][2] corrupted, bits corrected
};
cudaError_t launch_deint_rs(const DeintRsArgs &a, cudaStream_t st);

struct RsOnlyArgs {
  const uint8_t *rs_in; uint64_t npackets; const uint8_t *gf_exp, *gf_log;
  uint8_t *rts_out; int32_t *flags;
};
cudaError_t launch_rs_only(const RsOnlyArgs &a, cudaStream_t st);

struct DerandArgs {
  const uint8_t *rts; uint64_t npackets;
  const uint8_t *pattern;    // 1504 bytes
  int32_t pos_in;            // carried pattern position (multiple of 188)
  uint8_t *ts_out; uint64_t ts_cap;
  uint64_t *counts;          // device [4]: kept, dropped, pos_out, rs_errs_sum
  const int32_t *flags;      // RS flags for the error sum (may be null)
  uint32_t *scratch;         // [2 * npackets + 16 * (npackets / 1024 + 1)]: per-packet index and position, tile records
};
cudaError_t launch_derand(const DerandArgs &a, cudaStream_t st, int *launches);

}  // namespace ldvb
