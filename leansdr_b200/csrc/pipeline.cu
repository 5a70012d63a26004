// pipeline.cu -- handle, stream bookkeeping and the C ABI (include/leandvb_b200.h).
//
// One handle owns the device-side equivalent of the reference pipebufs between
// leandvb's input and p_tspackets (apps/leandvb.cc:204-596).  Each inter-stage
// stream is a flat device buffer "[items carried from the previous batch | items
// produced by this batch]": what a stage cannot consume yet stays at the front,
// exactly like unread items in a reference pipebuf (framework.h:153-159).
//
// Control decisions that the reference takes per run() call (filter retune,
// hypothesis switch, lock/unlock) are taken per batch on the host from small
// device-side result records; all sample/symbol/byte arithmetic runs in the
// kernels of k_*.cu.  There is no CPU fallback: without a CUDA device
// ldvb_create fails with LDVB_ENODEV.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/leandvb_b200.h"
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "kernels.h"
#include "tables.h"

using namespace ldvb;

namespace ldvb {
// defined in k_fec.cu
cudaError_t launch_deconv_carry(const DeconvArgs &a, uint64_t nsym, uint64_t *carry_out, cudaStream_t st);
}

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t alloc(size_t n) {
    bytes = n;
    return cudaMalloc(&p, n ? n : 16);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
  }
  template <class T> T *as() const { return static_cast<T *>(p); }
};

// Flat FIFO of `elem`-byte items on the device.
struct Stream {
  DevBuf buf;
  size_t elem = 1;
  uint64_t cap = 0;     // items
  uint64_t count = 0;   // items currently held (carry + new)
  uint64_t fresh = 0;   // items appended by the current batch (for taps)
  uint64_t head = 0;    // read position: items in front of it were consumed lazily (stream_consume_lazy)
  uint8_t *at(uint64_t i) const { return buf.as<uint8_t>() + (head + i) * elem; }
};

struct HypState { uint64_t reg = 0, acc = 0; int n_in = 0, n_out = 0; };

constexpr int kMeasGroup = 256;                 // cnr/spectrum measurements per launch

struct Tap {
  DevBuf buf;
  uint64_t bytes = 0;
};

struct ProfSpanRec { int id; cudaEvent_t a, b; };

}  // namespace

struct RingState;
// Page-locked mailbox of the receiver stage: everything the host reads back after the spans ran arrives with ONE
// stream synchronisation (seam plan, the carried loop state, the level check, the telemetry rows).
struct RxMail {
  uint64_t plan[16];
  RxState last;
  float power;
  uint32_t nm;
  float rows[4 * 4096];
};
struct ldvb_handle {
  ldvb_config cfg;
  cudaStream_t st = nullptr;
  std::string err;
  ldvb_meas meas;
  uint32_t launches = 0;

  // ---- derived configuration
  int bps_in = 2;
  float Fs_rx = 0;
  int decim = 1;
  std::vector<float> fir_coeffs, fir_shifted;
  float fir_current_freq = 0, fir_tol = 0.1f, fir_tap_mult = 1;
  int fir_n = 0;
  bool use_fir = false, use_decim = false, use_rot = false;
  Cstln cst;
  DeconvPolys dec;
  RxParams rxp;
  std::vector<float> rrc_coeffs;
  int readahead = 1;

  // ---- device tables
  DevBuf d_rrc;
  DevBuf d_pe16;             // slicer 1 (QPSK): phase_error column of the constellation table
  DevBuf d_cstln, d_trig, d_rot, d_taps, d_gfexp, d_gflog, d_derand, d_rotperm, d_twiddle;

  // ---- streams
  Stream s_raw;      // head part of the raw stream (carry + start of batch, or whole batch for push)
  Stream s_notched;  // cf32 after the notch (input of the front end when anf > 0)
  Stream s_pp;       // cf32 preprocessed = receiver input
  Stream s_sym;      // softsymbols
  Stream s_bytes;    // deconvolved bytes
  Stream s_mpeg;     // aligned bytes (with de-interleaver history)
  DevBuf d_rts, d_rsflags, d_ts, d_scratch, d_badwords, d_rs204;
  uint64_t ts_cap = 0;

  // ---- carry state
  NotchState notch;            // host mirror
  std::map<int, uint32_t> notch_table_of_bin;
  DevBuf d_notch_tables; uint32_t notch_tables_used = 0, notch_tables_cap = 0;
  float *notch_stage = nullptr; size_t notch_stage_bytes = 0; cudaEvent_t notch_stage_ev = nullptr;   // page-locked staging of new tables
  DevBuf d_notch_guess, d_notch_weights, d_notch_list;
  DevBuf d_notch_edge, d_notch_dump, d_notch_dumpblocks;   // k_notchfir.cu: segment edges, telemetry blocks
  bool rx_cold = true;                  // no FAST batch has settled the AGC of this stream yet
  RxMail *rx_mail = nullptr;            // page-locked
  bool notch_v2 = false, notch_fused = false;
  uint64_t notch_target_segs = 0;
  DevBuf d_notch_state, d_notch_epochs, d_notch_entry, d_notch_exit, d_notch_exact, d_notch_bins, d_notch_blocks;
  uint32_t rot_index = 0;
  RxState rx_state;            // host mirror of the exact/carried receiver state
  DevBuf d_rx_state, d_rx_info, d_rx_end, d_rx_head, d_rx_tail, d_rx_seams, d_rx_spans;
  DevBuf d_rx_off, d_rx_skip, d_rx_rot, d_rx_meas, d_rx_measn, d_rx_forced, d_rx_begin, d_rx_power, d_rx_settled;
  uint32_t rx_max_spans = 1;
  uint64_t rx_target_spans = 0;   // lanes of one full wave of k_rx (set at the first FAST launch)
  double rx_sym_per_sample = 1;
  HypState hyp[4];
  HypState hyp2[4];          // --fastlock: auxiliary registers in2/n_in2/n_out2 (dvb.h:303-306)
  // --hs: dvb_deconvol_sync_hard carry (dvb.h:662-672): the last 32 hard symbols (2 bits each, newest
  // in the LSBs; every alignment's shift registers are a remapping of them), vote phase, lock
  uint64_t hs_hist = 0; int hs_hist_valid = 0, hs_resync_phase = 0, hs_locked = 0;
  DevBuf d_hs_polar, d_hs_rect, d_hs_sincos, d_hs_errors, d_hs_lock, d_hs_state;
  int locked = 0, skip = 0;
  DevBuf d_deconv_carry;
  // Viterbi (viterbi_sync)
  Trellis trellis; VitSyncs vsyncs;
  DevBuf d_vit_pred, d_vit_us, d_vit_maps, d_vit_shifts, d_vit_state, d_vit_ctl;
  int vit_wave = 0;                            // CTAs of k_viterbi resident at once on this device
  DevBuf d_vit_entry, d_vit_exit, d_vit_aux;   // per time segment: entry / exit states, ctl, lists
  SyncState sync;
  DevBuf d_sync_state, d_sync_res;
  int derand_pos = 0;
  DevBuf d_counts;

  // ---- cnr_fft / spectrum (telemetry in front of the FIR)
  struct MeasUnit {
    bool on = false;
    int logn = 12;
    float kavg = 0.1f, bandwidth = 0;
    int64_t decimation = 1, phase = 0;
    uint64_t pos = 0;              // absolute index of the next unread sample
    DevBuf d_avg, d_have;
  } m_cnr, m_spec;
  DevBuf d_meas_carry[2], d_meas_points, d_meas_power, d_meas_sums, d_meas_rows;
  int meas_carry_sel = 0;
  uint64_t meas_carry_count = 0;   // samples kept from the previous batch (< 4096)
  uint64_t meas_abs_next = 0;      // absolute index of the next new sample
  std::vector<float> cnr_queue, spec_queue;
  // Telemetry results still on their way to the host: the band sums / averaged rows are copied
  // asynchronously into page-locked memory and turned into dB (glibc logf / log10f, like the
  // reference) by meas_finish(), which run_chain calls while the receiver kernel is running.
  struct MeasPending { bool is_cnr; int np, n; size_t off; };
  std::vector<MeasPending> meas_pending;
  float *meas_host = nullptr; size_t meas_host_cap = 0, meas_host_used = 0;
  cudaEvent_t meas_ev = nullptr;
  // rate_estimator<float> (generic.h:272-305) on the RS counts: accumulators and queued ratios
  std::vector<float> vber_queue;
  int64_t vber_num = 0, vber_den = 0;
  int vber_sample = 50000;
  std::vector<int32_t> vber_flags;

  // ---- time-sharded mode (ldvb_shard_*): what the front stage leaves for the back stage
  struct Shard {
    bool det_valid = false, valid = false;
    ldvb_shard desc;
    std::vector<NotchEpoch> epochs;      // local block units
    // geometry (absolute units unless noted)
    uint64_t R0 = 0, base_chunk = 0, own_begin = 0, own_end = 0, avail_end = 0, notch_b0 = 0;
    bool first = false;
    uint64_t settle_syms = 0;             // symbols the first chunk's settling pass put at the front of s_sym
    RxArgs a; RxStitchArgs sa; NotchApplyArgs na;
    const float2 *pp = nullptr;
  } shard;
  DevBuf d_edge_tail, d_edge_state, d_edge_seam;

  // ---- pipelined host input (ldvb_push): copy engine fills one staging buffer while
  // the chain works on the other
  static constexpr int kStages = 3;
  DevBuf d_stage[kStages];
  cudaStream_t copy_st = nullptr;
  cudaEvent_t copy_done[kStages] = {nullptr, nullptr, nullptr};
  uint64_t sub_batch = 0;

  // ---- async_push: the chain of every staged sub-batch runs on this thread, in order; the caller's thread only
  // copies.  While jobs are pending the handle's stream and state belong to the worker: every other entry point
  // waits for the queue to drain first (async_wait).  The TS queue is the only structure shared under way (qmu).
  struct Job { int stage; uint64_t n; };
  std::thread worker;
  std::mutex amu;                       // jobs, stage_busy, pending, worker_rc, stop
  std::condition_variable acv_job, acv_done;
  std::deque<Job> jobs;
  bool stage_busy[kStages] = {false, false, false};
  int pending = 0;
  int worker_rc = 0;
  bool worker_stop = false, worker_started = false;
  uint64_t stage_next = 0;
  double dbg_stage_wait_ms = 0, dbg_copy_wait_ms = 0, dbg_chain_ms = 0; uint64_t dbg_jobs = 0;   // LDVB_ASYNC_DEBUG
  std::mutex qmu;                       // ts_queue, ts_queue_cap, rd, wr, ready
  struct RingState *ring = nullptr;     // ldvb_ring_*: the NCCL transport of the time-sharded mode
  // telemetry as of the last finished job (amu): what ldvb_get_meas / ldvb_pull_{cnr,vber,spectrum} serve while
  // the worker owns the live copies
  ldvb_meas meas_pub;
  std::vector<float> cnr_pub, vber_pub, spec_pub;

  // ---- host-side TS queue for push/pull: page-locked, so the D2H of a sub-batch's packets is a
  // true asynchronous DMA that overlaps the next sub-batch (bytes [rd, wr) are unread)
  uint8_t *ts_queue = nullptr;
  size_t ts_queue_cap = 0, ts_queue_rd = 0, ts_queue_wr = 0;
  size_t ts_queue_ready = 0;            // bytes [rd, ready) have arrived (their copies are complete)

  // ---- per-kernel timing (ldvb_profile)
  bool profiling = false;
  bool own_stream = true;
  std::vector<std::string> prof_names;
  std::vector<uint32_t> prof_launches;
  std::vector<double> prof_ms;
  std::vector<ProfSpanRec> prof_pending;
  std::vector<cudaEvent_t> prof_free;

  // ---- taps
  Tap taps[9];
  std::vector<float> meas_log;   // {freq_tap, ss, mer} per measurement of the last batch
};

namespace {

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      char b_[256];                                                                   \
      snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      h->err = b_;                                                                    \
      return LDVB_ECUDA;                                                              \
    }                                                                                 \
  } while (0)

// Kernel launch with optional per-kernel CUDA-event timing (ldvb_profile).
struct ProfSpan { int id; cudaEvent_t a, b; };

int prof_begin(ldvb_handle *h, const char *name, ProfSpan *sp);
void prof_end(ldvb_handle *h, ProfSpan *sp);

#define KL(name, call)                                                                \
  do {                                                                                \
    ProfSpan sp_;                                                                     \
    const int prof_ = h->profiling ? prof_begin(h, name, &sp_) : 0;                   \
    cudaError_t e_ = (call);                                                          \
    if (prof_) prof_end(h, &sp_);                                                     \
    ++h->launches;                                                                    \
    if (e_ != cudaSuccess) {                                                          \
      char b_[256];                                                                   \
      snprintf(b_, sizeof b_, "%s:%d: kernel %s: %s", __FILE__, __LINE__, name, cudaGetErrorString(e_)); \
      h->err = b_;                                                                    \
      return LDVB_ECUDA;                                                              \
    }                                                                                 \
  } while (0)

struct NotchTableCache;
NotchTableCache *notch_table_cache();   // (defined with the notch stage)

// Waits for the background chain (async_push) and returns its error, if any.  While jobs are pending the
// handle's stream and state belong to the worker thread: every entry point except push / pull calls this first.
int async_wait(ldvb_handle *h) {
  if (!h->worker_started) return LDVB_OK;
  std::unique_lock<std::mutex> lk(h->amu);
  h->acv_done.wait(lk, [&] { return h->pending == 0; });
  return h->worker_rc;
}

int fail(ldvb_handle *h, int code, const char *msg) {
  h->err = msg;
  return code;
}

int prof_begin(ldvb_handle *h, const char *name, ProfSpan *sp) {
  int id = -1;
  for (size_t i = 0; i < h->prof_names.size(); ++i)
    if (h->prof_names[i] == name) { id = (int)i; break; }
  if (id < 0) {
    id = (int)h->prof_names.size();
    h->prof_names.push_back(name);
    h->prof_launches.push_back(0);
    h->prof_ms.push_back(0);
  }
  auto get = [&](cudaEvent_t *e) {
    if (!h->prof_free.empty()) { *e = h->prof_free.back(); h->prof_free.pop_back(); return true; }
    return cudaEventCreate(e) == cudaSuccess;
  };
  if (!get(&sp->a) || !get(&sp->b)) return 0;
  sp->id = id;
  cudaEventRecord(sp->a, h->st);
  return 1;
}

void prof_end(ldvb_handle *h, ProfSpan *sp) {
  cudaEventRecord(sp->b, h->st);
  h->prof_pending.push_back({sp->id, sp->a, sp->b});
}

void prof_harvest(ldvb_handle *h) {
  if (h->prof_pending.empty()) return;
  cudaStreamSynchronize(h->st);
  for (auto &r : h->prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      h->prof_ms[r.id] += ms;
      h->prof_launches[r.id] += 1;
    }
    h->prof_free.push_back(r.a);
    h->prof_free.push_back(r.b);
  }
  h->prof_pending.clear();
}

// Host wall-clock per stage (kernels + copies + synchronisations), reported next to the
// per-kernel device times as pseudo entries "wall:<stage>" while profiling is on.
struct WallTimer {
  ldvb_handle *h; const char *name;
  std::chrono::steady_clock::time_point t0;
  WallTimer(ldvb_handle *hh, const char *n) : h(hh), name(n), t0(std::chrono::steady_clock::now()) {}
  ~WallTimer() {
    if (!h->profiling) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    int id = -1;
    for (size_t i = 0; i < h->prof_names.size(); ++i) if (h->prof_names[i] == name) { id = (int)i; break; }
    if (id < 0) { id = (int)h->prof_names.size(); h->prof_names.push_back(name); h->prof_launches.push_back(0); h->prof_ms.push_back(0); }
    h->prof_ms[id] += ms; h->prof_launches[id] += 1;
  }
};

cudaError_t upload(DevBuf &b, const void *src, size_t n) {
  cudaError_t e = b.alloc(n);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(b.p, src, n, cudaMemcpyHostToDevice);
}

int stream_alloc(ldvb_handle *h, Stream &s, size_t elem, uint64_t cap) {
  s.elem = elem;
  s.cap = cap;
  s.count = 0;
  CK(s.buf.alloc((size_t)(cap + 64) * elem + 256));
  CK(cudaMemset(s.buf.p, 0, s.buf.bytes));
  return LDVB_OK;
}

// Drop `n` items from the front, moving the remainder to offset 0.
int stream_consume(ldvb_handle *h, Stream &s, uint64_t n, DevBuf &tmp) {
  if (n > s.count) return fail(h, LDVB_ESTATE, "stream underflow");
  const uint64_t left = s.count - n;
  if (n && left) {
    if (left <= n) {
      CK(cudaMemcpyAsync(s.at(0), s.at(n), left * s.elem, cudaMemcpyDeviceToDevice, h->st));
    } else {
      // Overlapping move: forward, in pieces staged through the scratch buffer.
      const uint64_t piece = tmp.bytes / s.elem;
      if (!piece) return fail(h, LDVB_EOVERFLOW, "no scratch for the carry");
      for (uint64_t off = 0; off < left; off += piece) {
        const uint64_t k = std::min(piece, left - off);
        CK(cudaMemcpyAsync(tmp.p, s.at(n + off), k * s.elem, cudaMemcpyDeviceToDevice, h->st));
        CK(cudaMemcpyAsync(s.at(off), tmp.p, k * s.elem, cudaMemcpyDeviceToDevice, h->st));
      }
    }
  }
  s.count = left;
  return LDVB_OK;
}

// Consume without moving the remainder: only the read position advances.  For streams that are drained in
// several passes per batch (the symbols, while mpeg_sync is searching); stream_compact brings the (then small)
// remainder back to the front of the buffer.
int stream_consume_lazy(ldvb_handle *h, Stream &s, uint64_t n) {
  if (n > s.count) return fail(h, LDVB_ESTATE, "stream underflow");
  s.head += n;
  s.count -= n;
  if (!s.count) s.head = 0;
  return LDVB_OK;
}

int stream_compact(ldvb_handle *h, Stream &s, DevBuf &tmp) {
  if (!s.head) return LDVB_OK;
  const uint64_t n = s.head, left = s.count;
  s.head = 0;
  s.count = left + n;                    // view the buffer from its start again, then drop the consumed items
  return stream_consume(h, s, n, tmp);
}

int tap_store(ldvb_handle *h, int which, const void *dev, uint64_t bytes) {
  if (!h->cfg.keep_taps) return LDVB_OK;
  Tap &t = h->taps[which];
  if (t.buf.bytes < bytes) {
    t.buf.release();
    CK(t.buf.alloc(bytes + bytes / 4 + 4096));
  }
  if (bytes) CK(cudaMemcpyAsync(t.buf.p, dev, bytes, cudaMemcpyDeviceToDevice, h->st));
  t.bytes = bytes;
  return LDVB_OK;
}

// ------------------------------------------------------------------ receiver

void rx_reset_state(ldvb_handle *h) {
  h->rx_cold = true;
  RxState &s = h->rx_state;
  memset(&s, 0, sizeof s);
  s.est_insp = 75.0f * 75.0f;  // sdr.h:727
  s.agc_gain = 1;
}

// Everything a fresh handle starts from (also ldvb_reset).
void reset_carry(ldvb_handle *h) {
  Stream *ss[] = {&h->s_raw, &h->s_notched, &h->s_pp, &h->s_sym, &h->s_bytes, &h->s_mpeg};
  for (Stream *s : ss) { s->count = 0; s->fresh = 0; s->head = 0; }
  memset(&h->notch, 0, sizeof h->notch);
  h->notch.gain = 1;
  for (int s = 0; s < kNotchMaxSlots; ++s) h->notch.slot[s].bin = -1;
  h->rot_index = 0;
  const float freqw = h->cfg.Ftune ? (h->cfg.Ftune / h->Fs_rx) * 65536 : 0.0f;
  rx_reset_state(h);
  h->rx_state.freqw = freqw;
  h->rx_state.freq_tap = freqw / 65536;
  if (h->cfg.hs) {   // fast_qpsk_receiver: integer loop state, no AGC (sdr.h:957-987)
    const long fw = h->cfg.Ftune ? (long)((h->cfg.Ftune / h->cfg.Fs) * 65536) : 0;
    h->rx_state.est_insp = 0; h->rx_state.agc_gain = 0;
    h->rx_state.freqw = (float)fw;
    h->rx_state.freq_tap = (float)fw / 65536;
  }
  if (h->use_fir && h->fir_current_freq != 0) {
    h->fir_shifted = shift_taps(h->fir_coeffs, 0);
    h->fir_current_freq = 0;
    cudaMemcpy(h->d_taps.p, h->fir_shifted.data(), h->fir_shifted.size() * 4, cudaMemcpyHostToDevice);
  }
  for (HypState &s : h->hyp) s = HypState();
  for (HypState &s : h->hyp2) s = HypState();
  h->locked = 0; h->skip = 0;
  if (h->d_vit_state.p) {   // viterbi_dec constructor: metrics 0, paths 0 (viterbi.h:133-145)
    cudaMemset(h->d_vit_state.p, 0, h->d_vit_state.bytes);
    cudaMemset(h->d_vit_ctl.p, 0, h->d_vit_ctl.bytes);
  }
  memset(&h->sync, 0, sizeof h->sync);
  h->sync.report_state = 1;
  h->sync.phase8 = -1;
  h->sync.fastlock = (h->cfg.fastlock || h->cfg.hs) ? 1 : 0;             // leandvb.cc:565, 862
  h->sync.resync_period = (h->cfg.hs && !h->cfg.fastlock) ? 32 : 1;      // dvb.h:729, leandvb.cc:553, 863
  h->hs_hist = 0; h->hs_hist_valid = 0; h->hs_resync_phase = 0; h->hs_locked = 0;
  h->derand_pos = 0;
  {   // (ldvb_pull may be running on a second thread of an async_push host)
    std::lock_guard<std::mutex> lk(h->qmu);
    h->ts_queue_rd = h->ts_queue_wr = h->ts_queue_ready = 0;
  }
  memset(&h->meas, 0, sizeof h->meas);
  for (ldvb_handle::MeasUnit *u : {&h->m_cnr, &h->m_spec}) {
    u->phase = 0; u->pos = 0;
    if (u->d_have.p) cudaMemset(u->d_have.p, 0, 4);
  }
  h->meas_carry_count = 0; h->meas_abs_next = 0;
  h->cnr_queue.clear(); h->spec_queue.clear();
  h->meas_pending.clear(); h->meas_host_used = 0;      // (callers have synchronised the stream)
  h->vber_queue.clear(); h->vber_num = h->vber_den = 0;
  h->vber_sample = std::max(50000, (int)(h->cfg.Fm / 2));   // leandvb.cc:585-587
  // The clears above went to the legacy default stream, which the handle's non-blocking stream does not
  // synchronise with: make them visible before anything is launched on h->st.
  cudaDeviceSynchronize();
}

void rx_setup(ldvb_handle *h) {
  const ldvb_config &c = h->cfg;
  RxParams &p = h->rxp;
  memset(&p, 0, sizeof p);
  for (int s = 0; s < h->cst.nsymbols && s < 256; ++s) {
    p.sym_re[s] = h->cst.sym_re[s];
    p.sym_im[s] = h->cst.sym_im[s];
  }
  p.nsymbols = h->cst.nsymbols;
  p.sampler = c.sampler;
  // set_omega (sdr.h:738-743) then update_freq_limits (sdr.h:755-770) with the
  // constellation already known (leandvb.cc:476-482).
  const float omega = h->Fs_rx / c.Fm;
  const float tol = 10e-6;
  const float max_omega = omega * (1 + tol);
  int n = 4;
  switch (h->cst.nsymbols) { case 2: n = 2; break; case 4: n = 4; break; case 8: n = 8; break;
                             case 16: n = 12; break; case 32: n = 16; break; default: n = 4; }
  float freqw = 0;
  if (c.Ftune) freqw = (c.Ftune / h->Fs_rx) * 65536;  // set_freq (sdr.h:745-749)
  // The constructor's set_freq(0)/set_omega(1) are overwritten by these calls;
  // note that set_freq() after set_omega() recomputes the limits around freqw.
  p.omega = omega;
  p.min_freqw = freqw - 65536 / max_omega / n / 2;
  p.max_freqw = freqw + 65536 / max_omega / n / 2;
  float pll_adjustment = 1.0f;
  if (c.viterbi) pll_adjustment /= 6;  // leandvb.cc:498-501
  p.freq_alpha = 0.04;                 // sdr.h:776-778
  p.freq_beta = 0.0012 / omega * pll_adjustment;
  p.gain_mu = 0.02 / (75.0f * 75.0f) * 2;
  p.kest = 0.01f;
  p.allow_drift = c.allow_drift;
  int md = (int)(h->Fs_rx / c.Finfo);  // decimation(Fs, Finfo), leandvb.cc:138-141, 502
  p.meas_decimation = (uint32_t)std::max(md, 1);
  rx_reset_state(h);
  h->rx_state.freqw = freqw;
  h->rx_state.freq_tap = freqw / 65536;
  h->readahead = (c.sampler == LDVB_SAMP_NEAREST) ? 0 : 1;
  if (c.sampler == LDVB_SAMP_RRC) {   // leandvb.cc:437-456
    int steps = 0;
    h->rrc_coeffs = design_rrc(h->Fs_rx, c.Fm, c.rolloff, c.rrc_rej, c.rrc_steps, &steps);
    p.rrc_n = (int)h->rrc_coeffs.size();
    p.rrc_sub = steps;
    h->readahead = p.rrc_n - 1;       // sdr.h:645
  }
  if (c.hs) {
    // fast_qpsk_receiver (sdr.h:977-1005): set_omega(Fs/Fm), set_freq(Ftune/Fs), limits +-65536/max_omega/8
    // in `signed long` arithmetic, freq_beta truncated to a long; all of them exact in a float.
    p.sampler = kRxSamplerHs;
    long fw = 0;
    if (c.Ftune) fw = (long)((c.Ftune / c.Fs) * 65536);           // set_freq: freqw = freq * 65536
    const long lo = (long)(fw - 65536 / max_omega / 8), hi = (long)(fw + 65536 / max_omega / 8);
    p.min_freqw = (float)lo; p.max_freqw = (float)hi;
    p.hs_freq_beta = (long long)(signed long)(0.0012 * 256 * 65536 / omega * 1.0f);
    int mdh = (int)(c.Fs / c.Finfo);
    p.meas_decimation = (uint32_t)std::max(mdh, 1);
    rx_reset_state(h);
    h->rx_state.est_insp = 0; h->rx_state.agc_gain = 0;
    h->rx_state.freqw = (float)fw;
    h->rx_state.freq_tap = (float)fw / 65536;
    h->readahead = 1;                                                // sdr.h:1010: chunk_size + 1
  }
}

}  // namespace

// =========================================================================== API

extern "C" {

int ldvb_abi_version(void) { return LDVB_ABI_VERSION; }

static int stage_init(ldvb_handle *h);   // (defined with ldvb_push)

const char *ldvb_strerror(int code) {
  switch (code) {
    case LDVB_OK: return "ok";
    case LDVB_EINVAL: return "invalid argument or unsupported configuration";
    case LDVB_ENOMEM: return "out of memory";
    case LDVB_ECUDA: return "CUDA error";
    case LDVB_ENODEV: return "no usable CUDA device (sm_100 required)";
    case LDVB_EOVERFLOW: return "batch larger than the handle was sized for";
    case LDVB_ESTATE: return "invalid call sequence / internal state";
    default: return "unknown error";
  }
}

const char *ldvb_last_error(const ldvb_handle *h) { return h ? h->err.c_str() : ""; }

void ldvb_config_default(ldvb_config *c) {
  memset(c, 0, sizeof *c);
  c->abi_version = LDVB_ABI_VERSION;
  c->input_format = LDVB_FMT_U8;
  c->float_scale = 1.0f;
  c->Fs = 2.4e6f;
  c->Fm = 2e6f;
  c->anf = 1;
  c->resample_rej = 10;
  c->sampler = LDVB_SAMP_LINEAR;
  c->rrc_rej = 10;
  c->rolloff = 0.35f;
  c->constellation = LDVB_CSTLN_QPSK;
  c->fec = LDVB_FEC12;
  c->Finfo = 5;
  c->rx_mode = LDVB_RX_EXACT;
  c->device = 0;
  c->max_batch = 1u << 22;
  c->spectrum = 1;   // leandvb.cc:333-343: always instantiated
  c->settle_chunks = 0;   // auto (512)
  c->seam_mode = 0;       // strict seams
}

int ldvb_destroy(ldvb_handle *h) {
  if (!h) return LDVB_OK;
  cudaSetDevice(h->cfg.device);
  if (h->ring) ldvb_ring_destroy(h);
  if (h->worker_started) {
    { std::unique_lock<std::mutex> lk(h->amu); h->acv_done.wait(lk, [&] { return h->pending == 0; }); h->worker_stop = true; }
    h->acv_job.notify_all();
    h->worker.join();
  }
  if (h->st) cudaStreamSynchronize(h->st);
  DevBuf *bufs[] = {&h->d_pe16, &h->d_rrc, &h->d_cstln, &h->d_trig, &h->d_rot, &h->d_taps, &h->d_gfexp, &h->d_gflog, &h->d_derand,
                    &h->d_rotperm, &h->d_twiddle, &h->s_raw.buf, &h->s_notched.buf, &h->s_pp.buf, &h->s_sym.buf,
                    &h->s_bytes.buf, &h->s_mpeg.buf, &h->d_rts, &h->d_rsflags, &h->d_ts, &h->d_scratch,
                    &h->d_badwords, &h->d_rs204, &h->d_notch_tables, &h->d_notch_state, &h->d_notch_epochs,
                    &h->d_notch_edge, &h->d_notch_dump, &h->d_notch_dumpblocks, &h->d_notch_entry, &h->d_notch_exit, &h->d_notch_guess, &h->d_notch_weights, &h->d_notch_list, &h->d_notch_exact, &h->d_notch_bins, &h->d_notch_blocks,
                    &h->d_rx_state, &h->d_rx_info, &h->d_rx_end, &h->d_rx_head, &h->d_rx_tail, &h->d_rx_seams,
                    &h->d_rx_spans, &h->d_rx_off, &h->d_rx_skip, &h->d_rx_rot, &h->d_rx_meas, &h->d_rx_measn,
                    &h->d_rx_forced, &h->d_rx_begin, &h->d_rx_power, &h->d_rx_settled, &h->d_deconv_carry, &h->d_vit_pred, &h->d_vit_us, &h->d_vit_maps, &h->d_vit_shifts, &h->d_vit_state, &h->d_vit_ctl, &h->d_vit_entry, &h->d_vit_exit, &h->d_vit_aux, &h->d_sync_state, &h->d_sync_res, &h->d_counts,
                    &h->d_edge_tail, &h->d_edge_state, &h->d_edge_seam, &h->d_hs_polar, &h->d_hs_rect, &h->d_hs_sincos, &h->d_hs_errors, &h->d_hs_lock, &h->d_hs_state,
                    &h->m_cnr.d_avg, &h->m_cnr.d_have, &h->m_spec.d_avg, &h->m_spec.d_have, &h->d_meas_carry[0], &h->d_meas_carry[1],
                    &h->d_meas_points, &h->d_meas_power, &h->d_meas_sums, &h->d_meas_rows};
  for (DevBuf *b : bufs) b->release();
  for (Tap &t : h->taps) t.buf.release();
  for (int i = 0; i < ldvb_handle::kStages; ++i) { h->d_stage[i].release(); if (h->copy_done[i]) cudaEventDestroy(h->copy_done[i]); }
  if (h->copy_st) cudaStreamDestroy(h->copy_st);
  if (h->ts_queue) cudaFreeHost(h->ts_queue);
  if (h->rx_mail) cudaFreeHost(h->rx_mail);
  if (h->notch_stage) cudaFreeHost(h->notch_stage);
  if (h->notch_stage_ev) cudaEventDestroy(h->notch_stage_ev);
  if (h->meas_host) cudaFreeHost(h->meas_host);
  if (h->meas_ev) cudaEventDestroy(h->meas_ev);
  for (auto &r : h->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto &e : h->prof_free) cudaEventDestroy(e);
  if (h->st && h->own_stream) cudaStreamDestroy(h->st);
  delete h;
  return LDVB_OK;
}

int ldvb_create(const ldvb_config *cfg, ldvb_handle **out) {
  if (!cfg || !out) return LDVB_EINVAL;
  *out = nullptr;
  if (cfg->abi_version != LDVB_ABI_VERSION) return LDVB_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) return LDVB_ENODEV;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return LDVB_ENODEV;
  if (prop.major < 10) return LDVB_ENODEV;  // kernels are built for sm_100a only
  ldvb_handle *h = new (std::nothrow) ldvb_handle();
  if (!h) return LDVB_ENOMEM;
  h->cfg = *cfg;
  memset(&h->meas, 0, sizeof h->meas);
  const ldvb_config &c = h->cfg;
  auto bail = [&](int code, const char *msg) {
    fprintf(stderr, "ldvb_create: %s\n", msg);
    ldvb_destroy(h);
    return code;
  };
  if (c.input_format < 0 || c.input_format > LDVB_FMT_F32) return bail(LDVB_EINVAL, "bad input_format");
  if (c.hs) {
    // run_highspeed (leandvb.cc:727-969): u8 input (:771-772), code rate 1/2 only (:845-846), QPSK by
    // construction; that graph has no notch, rotator, filter, CNR or spectrum blocks.
    if (c.input_format != LDVB_FMT_U8) return bail(LDVB_EINVAL, "--hs requires --u8");
    if (c.fec != LDVB_FEC12) return bail(LDVB_EINVAL, "--hs currently supports code rate 1/2 only");
    if (c.constellation != LDVB_CSTLN_QPSK || c.viterbi || c.resample || c.decim > 1 || c.Fderot != 0)
      return bail(LDVB_EINVAL, "--hs: QPSK without --viterbi/--resample/--decim/--derotate");
    h->cfg.anf = 0; h->cfg.cnr = 0; h->cfg.spectrum = 0; h->cfg.sampler = LDVB_SAMP_LINEAR; h->cfg.hard_metric = 0;
  }
  if (c.cnr && c.Fm / c.Fs > 0.25f) return bail(LDVB_EINVAL, "CNR estimator requires Fsampling > 4x Fsignal");   // sdr.h:1283-1284
  if (c.sampler < 0 || c.sampler > 2) return bail(LDVB_EINVAL, "bad sampler");
  if (c.anf < 0 || c.anf > kNotchMaxSlots) return bail(LDVB_EINVAL, "anf must be 0..4");
  if (!(c.Fs > 0) || !(c.Fm > 0) || c.max_batch == 0) return bail(LDVB_EINVAL, "bad rates or max_batch");
  if (cudaSetDevice(c.device) != cudaSuccess) return bail(LDVB_ENODEV, "cudaSetDevice failed");
  if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) return bail(LDVB_ECUDA, "stream");

  // ---- tables
  h->cst = make_cstln(c.constellation, c.fec, c.hard_metric != 0);
  if (h->cst.nsymbols == 0) return bail(LDVB_EINVAL, "constellation not supported");
  // FAST spans restart the loops from the carried state a few chunks early.  That reproduces the
  // decisions of constant-envelope constellations (the slicer of BPSK/QPSK/8PSK looks at the angle
  // only), but the AGC estimate (sdr.h:863-869) remembers ~100 chunks and the ring / grid
  // decisions of APSK and QAM depend on it: those constellations always run the exact receiver.
  if (h->cfg.rx_mode == LDVB_RX_FAST && h->cst.nsymbols > 8) h->cfg.rx_mode = LDVB_RX_EXACT;
  int fec = c.fec;
  if (c.viterbi && fec == LDVB_FEC23 && (h->cst.nsymbols == 4 || h->cst.nsymbols == 64)) fec = LDVB_FEC46;  // leandvb.cc:533-537
  if (!c.viterbi && !make_deconv(fec, &h->dec)) return bail(LDVB_EINVAL, "code rate not supported");
  if (c.viterbi) {
    if (!make_trellis(fec, &h->trellis)) return bail(LDVB_EINVAL, "code rate not supported by the Viterbi decoder");
    int bps = 0; while ((1 << bps) < h->cst.nsymbols) ++bps;
    if (h->trellis.bits_out % bps) return bail(LDVB_EINVAL, "code rate not suitable for this constellation");  // dvb.h:1247-1251
    h->vsyncs = make_vitsyncs(h->cst, h->trellis);
    if (h->vsyncs.nsyncs > 16) return bail(LDVB_EINVAL, "too many Viterbi hypotheses");
    std::vector<uint8_t> maps;
    for (auto &m : h->vsyncs.map) maps.insert(maps.end(), m.begin(), m.end());
    std::vector<int32_t> shifts(h->vsyncs.shift.begin(), h->vsyncs.shift.end());
    bool vok = upload(h->d_vit_pred, h->trellis.pred.data(), h->trellis.pred.size()) == cudaSuccess &&
               upload(h->d_vit_us, h->trellis.us.data(), h->trellis.us.size()) == cudaSuccess &&
               upload(h->d_vit_maps, maps.data(), maps.size()) == cudaSuccess &&
               upload(h->d_vit_shifts, shifts.data(), shifts.size() * 4) == cudaSuccess &&
               h->d_vit_state.alloc(sizeof(VitDecState) * h->vsyncs.nsyncs) == cudaSuccess &&
               h->d_vit_ctl.alloc(sizeof(VitCtl)) == cudaSuccess;
    if (!vok) return bail(LDVB_ECUDA, "viterbi tables");
  }
  std::vector<float> trig = make_trig16();
  uint8_t gexp[512], glog[256];
  make_rs_tables(gexp, glog);
  std::vector<uint8_t> derand = make_derand_pattern();
  std::vector<uint8_t> rotperm;
  for (auto &r : h->cst.rot) rotperm.insert(rotperm.end(), r.begin(), r.end());
  // FFT twiddles for the notch detector: omega_rev (dsp.h:70-76)
  std::vector<float> tw(2 * kNotchN);
  for (int i = 0; i < kNotchN; ++i) {
    float a = (float)(2.0 * M_PI * i / kNotchN);
    tw[2 * i] = cosf(a);
    tw[2 * i + 1] = -sinf(a);
  }
  bool ok = upload(h->d_cstln, h->cst.cells.data(), h->cst.cells.size() * sizeof(CstlnCell)) == cudaSuccess &&
            upload(h->d_trig, trig.data(), trig.size() * 4) == cudaSuccess &&
            upload(h->d_gfexp, gexp, 512) == cudaSuccess && upload(h->d_gflog, glog, 256) == cudaSuccess &&
            upload(h->d_derand, derand.data(), derand.size()) == cudaSuccess &&
            upload(h->d_rotperm, rotperm.data(), rotperm.size()) == cudaSuccess &&
            upload(h->d_twiddle, tw.data(), tw.size() * 4) == cudaSuccess;
  if (!ok) return bail(LDVB_ECUDA, "table upload failed");

  // ---- front end (leandvb.cc:310-399)
  h->bps_in = frontend_bytes_per_sample(c.input_format);
  float Fs = c.Fs;
  h->use_rot = (c.Fderot != 0);
  if (h->use_rot) {
    std::vector<float> lut = make_rotator_lut(-c.Fderot / c.Fs);
    if (upload(h->d_rot, lut.data(), lut.size() * 4) != cudaSuccess) return bail(LDVB_ECUDA, "rotator upload");
  }
  h->decim = 1;
  if (c.resample) {
    int d = 1;
    h->fir_coeffs = design_resampler(Fs, c.Fm, c.rolloff, c.resample_rej, c.decim, &d);
    h->fir_n = (int)h->fir_coeffs.size();
    h->decim = d;
    h->use_fir = true;
    Fs /= d;
    h->fir_tap_mult = 1.0f / d;                     // leandvb.cc:508
    h->fir_tol = (float)((double)(c.Fm / (Fs * d)) * 0.1);   // leandvb.cc:509 (Fs already divided; 0.1 is a double there)
    h->fir_shifted = shift_taps(h->fir_coeffs, 0);
    h->fir_current_freq = 0;
    if (upload(h->d_taps, h->fir_shifted.data(), h->fir_shifted.size() * 4) != cudaSuccess)
      return bail(LDVB_ECUDA, "taps upload");
  } else if (c.decim > 1) {
    h->decim = (int)c.decim;
    h->use_decim = true;
    Fs /= h->decim;
  }
  h->Fs_rx = Fs;
  rx_setup(h);
  {
    // QPSK with the soft metric (the headline configuration): the slicer's symbol and cost are computed and the
    // phase error comes from a 128 KB int16 column that k_rx keeps in shared memory (kernels.h: slicer 1).
    // LDVB_RX_SLICER=0 forces the cell-table gather (A/B measurements); other constellations always use it.
    const char *e = getenv("LDVB_RX_SLICER");
    const bool want = !(e && atoi(e) == 0);
    if (want && c.constellation == LDVB_CSTLN_QPSK && !c.hard_metric && !c.hs) {
      // Folded over Q: phase_error(I, -Q) == -phase_error(I, Q) holds for every cell of the host-built table
      // (the I reflection does not: 18 cells differ by one unit), so the column shrinks to 256 x 129 entries
      // [I & 0xff][|Q|], 66 KB instead of 128 KB of shared memory -- room for more resident warps.  The
      // identity is CHECKED here over all 65536 cells; if it ever failed the generic cell-table gather stays.
      std::vector<int16_t> pe((size_t)256 * kPeFoldPitch, 0);
      for (int ib = 0; ib < 256; ++ib)
        for (int q = 0; q <= 128; ++q) {
          const int qb = (q == 128) ? 0x80 : q;                    // |Q| = 128 only exists as Q = -128
          const int v = h->cst.cells[(size_t)ib * 256 + qb].phase_error;
          pe[(size_t)ib * kPeFoldPitch + q] = (int16_t)((q == 128) ? -v : v);
        }
      bool folds = true;
      for (int ib = 0; ib < 256 && folds; ++ib)
        for (int qb = 0; qb < 256; ++qb) {
          const int Q = (int)(int8_t)qb, aq = Q < 0 ? -Q : Q;
          const int v = pe[(size_t)ib * kPeFoldPitch + aq];
          if ((Q < 0 ? -v : v) != h->cst.cells[(size_t)ib * 256 + qb].phase_error) { folds = false; break; }
        }
      if (folds) {
        if (upload(h->d_pe16, pe.data(), pe.size() * 2) != cudaSuccess) return bail(LDVB_ECUDA, "pe16 upload");
        h->rxp.pe16 = h->d_pe16.as<int16_t>();
        h->rxp.slicer = 1;
      }
    }
  }
  if (c.hs) {
    const HsTables ht = make_hs_tables();
    if (upload(h->d_hs_polar, ht.polar.data(), ht.polar.size() * 4) != cudaSuccess ||
        upload(h->d_hs_rect, ht.rect.data(), ht.rect.size() * 2) != cudaSuccess ||
        upload(h->d_hs_sincos, ht.sincos.data(), ht.sincos.size() * 2) != cudaSuccess)
      return bail(LDVB_ECUDA, "--hs tables upload");
    h->rxp.hs_polar = h->d_hs_polar.as<uint32_t>();
    h->rxp.hs_rect = h->d_hs_rect.as<uint16_t>();
    h->rxp.hs_sincos = h->d_hs_sincos.as<uint16_t>();
  }
  if (c.sampler == LDVB_SAMP_RRC) {
    if ((h->rxp.rrc_n + h->rxp.rrc_sub - 1) / h->rxp.rrc_sub > 6) return bail(LDVB_EINVAL, "RRC sampler: more than 6 taps per symbol");
    if (upload(h->d_rrc, h->rrc_coeffs.data(), h->rrc_coeffs.size() * 4) != cudaSuccess) return bail(LDVB_ECUDA, "rrc upload");
    h->rxp.rrc_coeffs = h->d_rrc.as<float>();
  }
  // Host batches larger than this are pipelined (copy/compute overlap); push_sub_batch overrides.
  // Default: 192 MiB of input per sub-batch (24 Mi cf32 samples, 96 Mi complex<u8> samples): the copy of a
  // sub-batch has to outlast the chain on the previous one, whose cost per sample does not depend on the format.
  h->sub_batch = c.push_sub_batch > 0 ? (uint64_t)c.push_sub_batch : ((uint64_t)192 << 20) / h->bps_in;

  // ---- stream buffers
  const uint64_t M = c.max_batch;
  const uint64_t carry_raw = 4096 + (uint64_t)h->fir_n + h->decim + 64;
  const uint64_t pp_max = M / h->decim + 4096 + 512;
  const double sym_per_sample = 1.0 / std::max(1.0f, h->rxp.omega - 0.15f);
  const uint64_t sym_max = (uint64_t)(pp_max * sym_per_sample) + 4096;
  int bits_per_sym = 2;                                // QPSK / deconvol_sync: <= 2 bits per symbol out
  while ((1 << bits_per_sym) < h->cst.nsymbols) ++bits_per_sym;   // viterbi_sync: < log2(nsymbols) (code rate < 1)
  const uint64_t bytes_max = sym_max * bits_per_sym / 8 + 4096;
  const uint64_t pk_max = bytes_max / 204 + 16;
  int rc;
  if ((rc = stream_alloc(h, h->s_raw, h->bps_in, M + carry_raw))) return bail(rc, h->err.c_str());
  if (c.anf && (rc = stream_alloc(h, h->s_notched, 8, M + carry_raw))) return bail(rc, h->err.c_str());
  if ((rc = stream_alloc(h, h->s_pp, 8, pp_max + 1024))) return bail(rc, h->err.c_str());
  if ((rc = stream_alloc(h, h->s_sym, 4, sym_max + 4096))) return bail(rc, h->err.c_str());
  if ((rc = stream_alloc(h, h->s_bytes, 1, bytes_max + 4096))) return bail(rc, h->err.c_str());
  if ((rc = stream_alloc(h, h->s_mpeg, 1, bytes_max + 8192))) return bail(rc, h->err.c_str());
  h->ts_cap = pk_max;
  // Receiver spans (FAST): sized for the shortest span the auto rule can pick.
  h->rx_sym_per_sample = sym_per_sample;
  uint32_t nsp = 1;
  size_t span_bytes = 64;
  if (c.rx_mode == LDVB_RX_FAST) {
    const uint64_t nch = pp_max / kRxChunk + 1;
    const uint32_t smin = c.span_chunks ? c.span_chunks : 4;
    nsp = (uint32_t)(nch / smin + 2);
    const double cap_min = (smin + kRxVerifyChunks + 1) * kRxChunk * sym_per_sample + 64 + 4;   // (+4: span_cap is rounded up to 16-byte groups)
    span_bytes = (size_t)(std::max((double)nsp * cap_min, 1.25 * sym_max) * 4) + 4096;
  }
  h->rx_max_spans = nsp;
  bool aok =
      h->d_rts.alloc(pk_max * 188 + 256) == cudaSuccess && h->d_rsflags.alloc(pk_max * 8 + 64) == cudaSuccess &&
      h->d_ts.alloc(pk_max * 188 + 256) == cudaSuccess && h->d_scratch.alloc(std::max<uint64_t>(pk_max * 8 + pk_max / 16 + 8192, 1 << 20)) == cudaSuccess &&
      h->d_badwords.alloc(pk_max / 8 + 4096) == cudaSuccess && h->d_rs204.alloc(c.keep_taps ? pk_max * 204 + 256 : 16) == cudaSuccess &&
      h->d_notch_state.alloc(sizeof(NotchState)) == cudaSuccess && h->d_rx_state.alloc(sizeof(RxState)) == cudaSuccess &&
      h->d_rx_info.alloc(sizeof(RxSpanInfo) * nsp) == cudaSuccess && h->d_rx_end.alloc(sizeof(RxState) * nsp) == cudaSuccess &&
      h->d_rx_head.alloc(sizeof(RxSeamSym) * kRxSeamLog * (size_t)nsp) == cudaSuccess &&
      h->d_rx_tail.alloc(sizeof(RxSeamSym) * kRxSeamLog * (size_t)nsp) == cudaSuccess &&
      h->d_rx_seams.alloc(sizeof(RxSeam) * nsp) == cudaSuccess &&
      h->d_rx_spans.alloc(span_bytes) == cudaSuccess &&
      h->d_rx_off.alloc(8 * ((size_t)nsp + 1) + 16 * ((size_t)nsp / 1024 + 2)) == cudaSuccess && h->d_rx_skip.alloc(4 * (size_t)nsp) == cudaSuccess &&
      h->d_rx_rot.alloc(nsp) == cudaSuccess && h->d_rx_meas.alloc(16 * 4096) == cudaSuccess &&
      h->d_rx_measn.alloc(4) == cudaSuccess && h->d_rx_forced.alloc(sizeof(RxState)) == cudaSuccess &&
      h->d_rx_begin.alloc(sizeof(RxState) * nsp) == cudaSuccess && h->d_rx_power.alloc(16) == cudaSuccess &&
      h->d_rx_settled.alloc(sizeof(RxState) + sizeof(RxSpanInfo) + 64) == cudaSuccess &&
      h->d_deconv_carry.alloc(256) == cudaSuccess && h->d_sync_state.alloc(sizeof(SyncState)) == cudaSuccess &&
      h->d_sync_res.alloc(sizeof(SyncResult)) == cudaSuccess && h->d_counts.alloc(128) == cudaSuccess;
  if (!aok) return bail(LDVB_ENOMEM, "device allocation failed");
  // cnr_fft / spectrum (leandvb.cc:322-343)
  {
    const int dec1 = std::max((int)(c.Fs / 1.0f), 1);   // decimation(cfg.Fs, 1), leandvb.cc:138-141
    h->m_cnr.on = c.cnr != 0; h->m_cnr.logn = 12; h->m_cnr.kavg = 0.1f; h->m_cnr.bandwidth = c.Fm / c.Fs; h->m_cnr.decimation = dec1;
    h->m_spec.on = c.spectrum != 0; h->m_spec.logn = 10; h->m_spec.kavg = 0.5f; h->m_spec.bandwidth = 0; h->m_spec.decimation = dec1;
    bool mok = true;
    for (ldvb_handle::MeasUnit *u : {&h->m_cnr, &h->m_spec})
      if (u->on) mok = mok && u->d_avg.alloc(4096 * 4) == cudaSuccess && u->d_have.alloc(4) == cudaSuccess &&
                       cudaMemset(u->d_have.p, 0, 4) == cudaSuccess;
    if (h->m_cnr.on || h->m_spec.on)
      mok = mok && h->d_meas_carry[0].alloc(4096 * 8) == cudaSuccess && h->d_meas_carry[1].alloc(4096 * 8) == cudaSuccess &&
            h->d_meas_points.alloc(kMeasGroup * 8) == cudaSuccess && h->d_meas_power.alloc((size_t)kMeasGroup * 4096 * 4) == cudaSuccess &&
            h->d_meas_sums.alloc(kMeasGroup * 12) == cudaSuccess && h->d_meas_rows.alloc((size_t)kMeasGroup * 1024 * 4) == cudaSuccess;
    if (!mok) return bail(LDVB_ENOMEM, "telemetry allocation failed");
  }
  // Notch
  memset(&h->notch, 0, sizeof h->notch);
  h->notch.gain = 1;
  for (int s = 0; s < kNotchMaxSlots; ++s) h->notch.slot[s].bin = -1;
  if (c.anf) {
    const uint64_t nblk = M / kNotchN + 2;
    h->notch_tables_cap = 4096;        // 128 MB: a flat spectrum visits a new bin at almost every detect point; 4096 = all of them
    bool nok = h->d_notch_tables.alloc((size_t)h->notch_tables_cap * kNotchN * 8) == cudaSuccess &&
               h->d_notch_epochs.alloc(sizeof(NotchEpoch) * (nblk / 1024 + 4)) == cudaSuccess &&
               h->d_notch_entry.alloc(8 * kNotchMaxSlots * (nblk + 1)) == cudaSuccess &&
               h->d_notch_exit.alloc(8 * kNotchMaxSlots * (nblk + 1)) == cudaSuccess &&
               h->d_notch_exact.alloc(nblk + 1) == cudaSuccess &&
               h->d_notch_guess.alloc(8 * kNotchMaxSlots * (nblk + 1)) == cudaSuccess &&
               h->d_notch_list.alloc(4 * (nblk + 1)) == cudaSuccess &&
               h->d_notch_bins.alloc(4 * kNotchMaxSlots * (nblk / 1024 + 4)) == cudaSuccess &&
               h->d_notch_blocks.alloc(8 * (nblk / 1024 + 4)) == cudaSuccess;
    if (!nok) return bail(LDVB_ENOMEM, "notch allocation failed");
    {  // (1-k)^m, m = 0..8191, for the start-state guess
      std::vector<float> w(8192);
      const double c1 = (double)(1.0f - 0.002f);
      for (int m = 0; m < 8192; ++m) w[m] = (float)pow(c1, (double)m);
      if (upload(h->d_notch_weights, w.data(), w.size() * 4) != cudaSuccess) return bail(LDVB_ECUDA, "notch weights");
    }
    (void)notch_table_cache();      // starts the background builders of the process-wide table cache (first handle only)
    // Table 0 is all zeros: slots that never detected (bin -1) use it (sdr.h:57-63).
    cudaMemset(h->d_notch_tables.p, 0, (size_t)kNotchN * 8);
    h->notch_tables_used = 1;
    // Kernel choice.  Default: the warp-specialised kernel of k_notchfir.cu, which -- when the low-pass that follows has
    // decimation 1, at most kFirFuseMaxTaps taps and no rotator in front -- applies the FIR on its store path, so that the
    // notched stream never reaches HBM (LDVB_NOTCH_FUSE=0: plain notch, then k_frontend).  Measured at the bench size
    // (profiles/r02_*): 1.20 ms + 0.06 ms of start-state sums against 0.90 + 0.32 + 0.17 ms for the lane-per-segment
    // kernel, k_frontend and the sums of 1-block segments, and 2.5 GB instead of 7.1 GB of DRAM traffic.
    // LDVB_NOTCH_V2=0 selects the lane-per-segment kernel of k_notch.cu (round 1); time-sharded handles always run it.
    const char *e2 = getenv("LDVB_NOTCH_V2"), *ef = getenv("LDVB_NOTCH_FUSE");
    h->notch_v2 = !(e2 && atoi(e2) == 0);
    h->notch_fused = h->notch_v2 && !(ef && atoi(ef) == 0) && h->use_fir && h->decim == 1 && !h->use_rot &&
                     h->fir_n >= 2 && h->fir_n <= kFirFuseMaxTaps;
  }
  reset_carry(h);
  if (c.async_push) {
    // A streaming host: the staging buffers and the page-locked packet queue are set up here, not inside the first
    // push (three cudaMallocs of up to 1 GiB and a cudaHostAlloc took 0.1-0.6 s of the first scheduler step,
    // depending on the box: bench.py e2e_runnable).
    int rcs = stage_init(h);
    if (rcs) return bail(rcs, h->err.c_str());
    const size_t cap = 2 * (size_t)h->ts_cap * 188;
    if (cudaHostAlloc((void **)&h->ts_queue, cap, cudaHostAllocDefault) != cudaSuccess) return bail(LDVB_ENOMEM, "TS queue");
    h->ts_queue_cap = cap;
  }
  *out = h;
  return LDVB_OK;
}

int ldvb_reset(ldvb_handle *h) {
  if (!h) return LDVB_EINVAL;
  async_wait(h);                         // (an error of the background chain does not outlive a reset)
  if (h->worker_started) { std::lock_guard<std::mutex> lk(h->amu); h->worker_rc = 0; }
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return fail(h, LDVB_ECUDA, "cudaSetDevice");
  CK(cudaStreamSynchronize(h->st));
  reset_carry(h);
  return LDVB_OK;
}

int ldvb_set_stream(ldvb_handle *h, void *cuda_stream) {
  if (!h) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  CK(cudaStreamSynchronize(h->st));
  if (h->own_stream && h->st) cudaStreamDestroy(h->st);
  h->st = static_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  return LDVB_OK;
}

int ldvb_profile(ldvb_handle *h, int enable) {
  if (!h) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  prof_harvest(h);
  h->profiling = enable != 0;
  if (enable) {
    std::fill(h->prof_ms.begin(), h->prof_ms.end(), 0.0);
    std::fill(h->prof_launches.begin(), h->prof_launches.end(), 0u);
  }
  return LDVB_OK;
}

int ldvb_get_profile(ldvb_handle *h, ldvb_kernel_stat *stats, int cap, int *n) {
  if (!h || !n) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  prof_harvest(h);
  int k = 0;
  for (size_t i = 0; i < h->prof_names.size() && k < cap; ++i) {
    if (!h->prof_launches[i]) continue;
    if (stats) {
      memset(&stats[k], 0, sizeof stats[k]);
      snprintf(stats[k].name, sizeof stats[k].name, "%s", h->prof_names[i].c_str());
      stats[k].launches = h->prof_launches[i];
      stats[k].ms_total = (float)h->prof_ms[i];
    }
    ++k;
  }
  *n = k;
  return LDVB_OK;
}

}  // extern "C"

namespace {

int meas_finish(ldvb_handle *h);
int meas_host_reserve(ldvb_handle *h, size_t need, float **dst);
int fir_retune(ldvb_handle *h, int *real_taps);
void meas_plan(ldvb_handle::MeasUnit &u, uint64_t abs0, uint64_t avail, std::vector<uint64_t> &points);
int meas_launch(ldvb_handle *h, ldvb_handle::MeasUnit &u, const MeasSrc &src, const std::vector<uint64_t> &points);


// ------------------------------------------------------------------------- notch

// Process-wide host cache of the 4096 possible expj tables (they depend on the bin only).  The first handle with a notch
// starts a few builder threads that fill it in the background (~0.5 core-seconds in all); a handle that needs a bin takes
// the finished table from here (a 32 KB copy) and only computes it itself when the builders have not got there yet.
// Without this a stream with a flat spectrum -- a new bin at almost every detect point -- paid ~130 us of libm calls per
// bin in front of every batch, with the GPU idle (0.35 ms per 128 M-sample step even on 8 threads).
struct NotchTableCache {
  std::vector<float> data;                       // [4096 bins][4096][2]
  std::unique_ptr<std::atomic<uint8_t>[]> ready; // 0: not built, 1: built
  std::vector<std::thread> builders;
  std::atomic<bool> stop{false};
  static void build(float *t, int bin) {
    for (int i = 0; i < kNotchN; ++i) {
      float a = (float)(2 * M_PI * bin * i / kNotchN);
      t[2 * i] = cosf(a);
      t[2 * i + 1] = sinf(a);
    }
  }
  NotchTableCache() : data((size_t)kNotchN * kNotchN * 2), ready(new std::atomic<uint8_t>[kNotchN]) {
    for (int b = 0; b < kNotchN; ++b) ready[b].store(0);
    const int nth = 4;
    for (int t = 0; t < nth; ++t)
      builders.emplace_back([this, t, nth] {
        for (int b = t; b < kNotchN && !stop.load(std::memory_order_relaxed); b += nth) {
          build(data.data() + (size_t)b * kNotchN * 2, b);
          ready[b].store(1, std::memory_order_release);
        }
      });
  }
  ~NotchTableCache() {
    stop.store(true);
    for (auto &x : builders) x.join();
  }
  // Copies the table of `bin` to dst; false when it is not built yet.
  bool get(int bin, float *dst) const {
    if (bin < 0 || bin >= kNotchN || !ready[bin].load(std::memory_order_acquire)) return false;
    memcpy(dst, data.data() + (size_t)bin * kNotchN * 2, (size_t)kNotchN * 8);
    return true;
  }
};
NotchTableCache *notch_table_cache() {
  static NotchTableCache cache;     // (constructed on first use, joined at process exit)
  return &cache;
}

// expj tables, built like auto_notch::detect() (sdr.h:104-108) with glibc's cosf / sinf on the host (bit parity).
// A table costs ~130 us of one core (2 x 4096 libm calls); a stream with a flat spectrum moves its notch to a new bin
// at almost every detect point (30 per 128 M samples), so all the new bins of a batch are built together on a few
// threads into page-locked memory and uploaded asynchronously.  The cache holds notch_tables_cap tables.
int notch_tables_prepare(ldvb_handle *h, const std::vector<int> &bins) {
  std::vector<int> todo;
  auto collect = [&] {
    todo.clear();
    for (int b : bins)
      if (b >= 0 && !h->notch_table_of_bin.count(b) && std::find(todo.begin(), todo.end(), b) == todo.end()) todo.push_back(b);
  };
  collect();
  if (todo.empty()) return LDVB_OK;
  if (h->notch_tables_used + todo.size() > h->notch_tables_cap) {
    // Recycle: forget everything except the zero table; every bin this batch needs is rebuilt below.
    h->notch_table_of_bin.clear();
    h->notch_tables_used = 1;
    collect();
    if (1 + todo.size() > h->notch_tables_cap) return fail(h, LDVB_EOVERFLOW, "more notch bins in one batch than the table cache holds");
  }
  const size_t tbytes = (size_t)kNotchN * 8;
  if (h->notch_stage_ev) CK(cudaEventSynchronize(h->notch_stage_ev));          // the previous upload has left the staging area
  else CK(cudaEventCreateWithFlags(&h->notch_stage_ev, cudaEventDisableTiming));
  if (todo.size() * tbytes > h->notch_stage_bytes) {
    if (h->notch_stage) cudaFreeHost(h->notch_stage);
    h->notch_stage = nullptr; h->notch_stage_bytes = 0;
    const size_t want = std::max<size_t>(todo.size(), 32) * tbytes;
    if (cudaHostAlloc((void **)&h->notch_stage, want, cudaHostAllocDefault) != cudaSuccess) return fail(h, LDVB_ENOMEM, "notch table staging");
    h->notch_stage_bytes = want;
  }
  // from the process-wide cache where the background builders have got there, else computed here (on a few threads)
  NotchTableCache *cache = notch_table_cache();
  std::vector<size_t> missing;
  for (size_t k = 0; k < todo.size(); ++k)
    if (!cache->get(todo[k], h->notch_stage + k * (size_t)kNotchN * 2)) missing.push_back(k);
  auto build = [&](size_t k) { NotchTableCache::build(h->notch_stage + k * (size_t)kNotchN * 2, todo[k]); };
  const size_t nth = std::min<size_t>(8, missing.size());
  if (nth <= 1) {
    for (size_t k : missing) build(k);
  } else {
    std::vector<std::thread> th;
    for (size_t t = 0; t < nth; ++t)
      th.emplace_back([&, t] { for (size_t q = t; q < missing.size(); q += nth) build(missing[q]); });
    for (auto &x : th) x.join();
  }
  for (size_t k = 0; k < todo.size(); ++k) {
    const uint32_t idx = h->notch_tables_used++;
    CK(cudaMemcpyAsync(h->d_notch_tables.as<uint8_t>() + (size_t)idx * tbytes, h->notch_stage + k * (size_t)kNotchN * 2, tbytes,
                       cudaMemcpyHostToDevice, h->st));
    h->notch_table_of_bin[todo[k]] = idx;
  }
  CK(cudaEventRecord(h->notch_stage_ev, h->st));
  return LDVB_OK;
}

int notch_table_for_bin(ldvb_handle *h, int bin, uint32_t *index) {
  if (bin < 0) { *index = 0; return LDVB_OK; }
  auto it = h->notch_table_of_bin.find(bin);
  if (it == h->notch_table_of_bin.end()) {
    int rc = notch_tables_prepare(h, std::vector<int>{bin});
    if (rc) return rc;
    it = h->notch_table_of_bin.find(bin);
  }
  *index = it->second;
  return LDVB_OK;
}

// Verifies entry(j) == exit(j-1) for every segment (device-side count first) and re-runs the
// segments that did not merge, all of them in one launch per round.  exitv receives the exit
// states (at least the last one).
int notch_verify_repair(ldvb_handle *h, NotchApplyArgs &a, std::vector<float2> &exitv_out, NotchFirArgs *fa = nullptr) {
  const ldvb_config &c = h->cfg;
  // Device-side check first: the host only reads a counter and the last exit state.
  CK(cudaMemsetAsync(h->d_counts.p, 0, 4, h->st));
  KL("notch_verify", launch_notch_verify(a, h->d_counts.as<uint32_t>(), h->st));
  uint32_t nfail = 0;
  std::vector<float2> entry;
  std::vector<float2> &exitv = exitv_out;
  exitv.assign((size_t)a.nsegs * kNotchMaxSlots, make_float2(0.f, 0.f));
  CK(cudaMemcpyAsync(&nfail, h->d_counts.p, 4, cudaMemcpyDeviceToHost, h->st));
  CK(cudaMemcpyAsync(&exitv[(size_t)(a.nsegs - 1) * kNotchMaxSlots], a.seg_exit + (size_t)(a.nsegs - 1) * kNotchMaxSlots,
                     8 * kNotchMaxSlots, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (nfail) {
  entry.resize((size_t)a.nsegs * kNotchMaxSlots);
  std::vector<uint8_t> exact(a.nsegs);
  CK(cudaMemcpyAsync(entry.data(), a.seg_entry, entry.size() * 8, cudaMemcpyDeviceToHost, h->st));
  CK(cudaMemcpyAsync(exact.data(), a.seg_exact, exact.size(), cudaMemcpyDeviceToHost, h->st));
  for (int round = 0; round < 1 << 20; ++round) {
    CK(cudaMemcpyAsync(exitv.data(), a.seg_exit, exitv.size() * 8, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    std::vector<uint32_t> todo;
    bool prev_failed = false;
    for (uint32_t j = 1; j < a.nsegs; ++j) {
      bool same = exact[j] != 0;
      if (!same) same = memcmp(&entry[(size_t)j * kNotchMaxSlots], &exitv[(size_t)(j - 1) * kNotchMaxSlots], 8 * (size_t)c.anf) == 0;
      if (!same && !prev_failed) todo.push_back(j);
      prev_failed = !same;
    }
    if (todo.empty()) break;
    CK(cudaMemcpyAsync(h->d_notch_list.p, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, h->st));
    if (fa) { fa->n = a; KL("notch_fir", launch_notch_fir(*fa, h->d_notch_list.as<uint32_t>(), (uint32_t)todo.size(), nullptr, h->st)); }
    else KL("notch_apply", launch_notch_apply(a, h->d_notch_list.as<uint32_t>(), (uint32_t)todo.size(), nullptr, h->st));
    h->meas.notch_repaired += (uint32_t)todo.size();
    for (uint32_t j : todo)   // by construction the repaired segment entered with exit(j-1)
      memcpy(&entry[(size_t)j * kNotchMaxSlots], &exitv[(size_t)(j - 1) * kNotchMaxSlots], 8 * kNotchMaxSlots);
  }
  }
  return LDVB_OK;
}

int run_notch(ldvb_handle *h, const RawSrc &src, uint64_t avail, uint64_t *consumed) {
  const ldvb_config &c = h->cfg;
  const uint64_t nblocks = avail / kNotchN;
  *consumed = nblocks * kNotchN;
  h->s_notched.fresh = 0;
  if (!nblocks) return LDVB_OK;
  // v2: the warp-specialised kernel (k_notchfir.cu); fused: the low-pass is applied on its store path and the
  // notched stream stays on the chip (s_notched only holds the fir_n samples carried between batches).
  const bool v2 = h->notch_v2, fused = h->notch_fused;
  if (!fused && h->s_notched.count + nblocks * kNotchN > h->s_notched.cap) return fail(h, LDVB_EOVERFLOW, "notched stream overflow");
  const int fmt = c.input_format;
  // Detect points (sdr.h:64-71): phase advances by 4096 before the test.
  std::vector<uint64_t> dblocks;
  {
    int64_t phase = h->notch.phase;
    for (uint64_t b = 0; b < nblocks; ++b) {
      phase += kNotchN;
      if (phase >= 1024 * 4096) { phase -= 1024 * 4096; dblocks.push_back(b); }
    }
    h->notch.phase = (int32_t)phase;
  }
  std::vector<NotchEpoch> epochs;
  {
    NotchEpoch e0;
    memset(&e0, 0, sizeof e0);
    e0.first_block = 0;
    for (int s = 0; s < c.anf; ++s) {
      e0.bin[s] = h->notch.slot[s].bin;
      int rc = notch_table_for_bin(h, e0.bin[s], &e0.table_index[s]);
      if (rc) return rc;
    }
    epochs.push_back(e0);
  }
  if (!dblocks.empty()) {
    NotchDetectArgs d;
    d.src = src; d.fmt = fmt; d.scale = c.float_scale;
    CK(cudaMemcpyAsync(h->d_notch_blocks.p, dblocks.data(), dblocks.size() * 8, cudaMemcpyHostToDevice, h->st));
    d.block_index = h->d_notch_blocks.as<uint64_t>();
    d.ndetect = (int)dblocks.size(); d.nslots = c.anf;
    d.twiddle_rev = h->d_twiddle.as<float2>();
    d.bins_out = h->d_notch_bins.as<int32_t>();
    KL("notch_detect", launch_notch_detect(d, h->st));
    std::vector<int32_t> bins(dblocks.size() * c.anf);
    CK(cudaMemcpyAsync(bins.data(), h->d_notch_bins.p, bins.size() * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    {  // every bin this batch will use: built together (see notch_tables_prepare)
      std::vector<int> need(bins.begin(), bins.end());
      for (int s = 0; s < c.anf; ++s) need.push_back(h->notch.slot[s].bin);
      int rcp = notch_tables_prepare(h, need);
      if (rcp) return rcp;
      for (int s = 0; s < c.anf; ++s) {                    // (a recycle may have moved the tables of the epoch in force)
        int rc2 = notch_table_for_bin(h, epochs[0].bin[s], &epochs[0].table_index[s]);
        if (rc2) return rc2;
      }
    }
    int cur[kNotchMaxSlots];
    for (int s = 0; s < c.anf; ++s) cur[s] = h->notch.slot[s].bin;
    for (size_t k = 0; k < dblocks.size(); ++k) {
      bool changed = false;
      NotchEpoch e = epochs.back();
      for (int s = 0; s < kNotchMaxSlots; ++s) e.reset[s] = 0;
      for (int s = 0; s < c.anf; ++s) {
        const int nb = bins[k * c.anf + s];
        if (nb != cur[s]) {  // sdr.h:97-109: new peak -> estimate reset, table rebuilt
          changed = true;
          cur[s] = nb;
          e.bin[s] = nb;
          e.reset[s] = 1;
          int rc = notch_table_for_bin(h, nb, &e.table_index[s]);
          if (rc) return rc;
        }
      }
      if (changed) {
        e.first_block = dblocks[k];
        if (e.first_block == 0) epochs[0] = e; else epochs.push_back(e);
      }
    }
    for (int s = 0; s < c.anf; ++s) h->notch.slot[s].bin = cur[s];
  }
  CK(cudaMemcpyAsync(h->d_notch_epochs.p, epochs.data(), epochs.size() * sizeof(NotchEpoch), cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(h->d_notch_state.p, &h->notch, sizeof(NotchState), cudaMemcpyHostToDevice, h->st));

  NotchApplyArgs a;
  a.src = src; a.fmt = fmt; a.scale = c.float_scale;
  a.out = reinterpret_cast<float2 *>(h->s_notched.at(h->s_notched.count));
  a.nblocks = nblocks; a.nslots = c.anf;
  a.block0 = 0; a.first_exact = 1;
  a.k = 0.002f; a.gain = h->notch.gain;
  a.w_block = (float)pow((double)(1.0f - 0.002f), (double)kNotchN);
  a.expj_tables = h->d_notch_tables.as<float2>();
  a.epochs = h->d_notch_epochs.as<NotchEpoch>();
  a.nepochs = (int)epochs.size();
  // Segment size: as many concurrent segments as the stream allows (the kernel is
  // latency bound per lane), 4 blocks (16 Ki samples) of warm-up: from a zero
  // estimate the float trajectories merge bit for bit after 4..10 Ki samples
  // (0.998^n decay below one ulp), measured in DESIGN.md.
  a.seg_blocks = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, (nblocks + 32767) / 32768));
  if (v2) {
    // One CTA = kNotchFirRows segments; as many segments as fit in one wave (two CTAs per SM with one slot, else one): the
    // chain warp needs ~8 cycles per sample, so longer segments cost little and dilute the warm-up blocks.
    if (!h->notch_target_segs) {
      int dev = 0, sms = 148;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      h->notch_target_segs = (uint64_t)kNotchFirRows * sms * (c.anf == 1 ? 2 : 1);
    }
    a.seg_blocks = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, (nblocks + h->notch_target_segs - 1) / h->notch_target_segs));
  }
  a.warm_blocks = 2;   // exact blocks after the parallel weighted-sum guess (merge: median 1 Ki, max ~3 Ki samples)
  {  // tuning knobs (experiments): LDVB_NOTCH_SEG = blocks per segment, LDVB_NOTCH_WARM = warm-up blocks
    static const int seg_env = [] { const char *e = getenv("LDVB_NOTCH_SEG"); return e ? atoi(e) : 0; }();
    static const int warm_env = [] { const char *e = getenv("LDVB_NOTCH_WARM"); return e ? atoi(e) : 0; }();
    if (seg_env > 0) a.seg_blocks = (uint32_t)seg_env;
    if (warm_env > 0) a.warm_blocks = (uint32_t)warm_env;
  }
  a.nsegs = (uint32_t)((nblocks + a.seg_blocks - 1) / a.seg_blocks);
  a.state_in = h->d_notch_state.as<NotchState>();
  a.seg_entry = h->d_notch_entry.as<float2>();
  a.seg_exit = h->d_notch_exit.as<float2>();
  a.seg_exact = h->d_notch_exact.as<uint8_t>();
  KL("notch_guess", launch_notch_guess(a, h->d_notch_guess.as<float2>(), h->d_notch_weights.as<float>(), h->st));
  NotchFirArgs fa;
  memset(&fa, 0, sizeof fa);
  uint64_t ycount = 0;
  std::vector<uint64_t> pts_cnr, pts_spec;
  if (v2) {
    if (fused) {
      int real = 1;
      { int rcr = fir_retune(h, &real); if (rcr) return rcr; }
      const uint64_t carry = h->s_notched.count;                       // 0 (first batch) or fir_n
      if (carry != 0 && carry != (uint64_t)h->fir_n) return fail(h, LDVB_ESTATE, "notched carry");
      ycount = carry + nblocks * kNotchN - (uint64_t)h->fir_n;         // dsp.h:246-247, decimation 1
      if (h->s_pp.count + ycount > h->s_pp.cap) return fail(h, LDVB_EOVERFLOW, "preprocessed stream overflow");
      if ((size_t)a.nsegs * kNotchEdge * 8 > h->d_notch_edge.bytes) {
        h->d_notch_edge.release();
        CK(h->d_notch_edge.alloc((size_t)a.nsegs * kNotchEdge * 8 + 65536));
      }
      fa.fir_n = h->fir_n; fa.real_taps = real; fa.taps = h->d_taps.as<float2>();
      fa.y = reinterpret_cast<float2 *>(h->s_pp.at(h->s_pp.count));
      fa.carry = (uint32_t)carry;
      fa.carry_in = reinterpret_cast<const float2 *>(h->s_notched.at(0));
      fa.carry_out = reinterpret_cast<float2 *>(h->s_notched.at(0));
      fa.edge = h->d_notch_edge.as<float2>();
      // cnr_fft / spectrum look at the notched stream (leandvb.cc:296-343): the blocks they will measure are
      // known in advance (sdr.h:1294-1302, 1362-1370) and are the only ones written out.
      if (h->m_cnr.on || h->m_spec.on) {
        if (h->meas_carry_count) return fail(h, LDVB_ESTATE, "telemetry carry in fused mode");
        const uint64_t abs0 = h->meas_abs_next;
        if (h->m_cnr.on) meas_plan(h->m_cnr, abs0, nblocks * kNotchN, pts_cnr);
        if (h->m_spec.on) meas_plan(h->m_spec, abs0, nblocks * kNotchN, pts_spec);
        std::vector<uint64_t> blocks;
        for (uint64_t pnt : pts_cnr) blocks.push_back(pnt / kNotchN);
        for (uint64_t pnt : pts_spec) blocks.push_back(pnt / kNotchN);
        std::sort(blocks.begin(), blocks.end());
        blocks.erase(std::unique(blocks.begin(), blocks.end()), blocks.end());
        if (!blocks.empty()) {
          if (blocks.size() * 8 > h->d_notch_dumpblocks.bytes) { h->d_notch_dumpblocks.release(); CK(h->d_notch_dumpblocks.alloc(blocks.size() * 8 + 1024)); }
          if (blocks.size() * kNotchN * 8 > h->d_notch_dump.bytes) { h->d_notch_dump.release(); CK(h->d_notch_dump.alloc((blocks.size() + 16) * kNotchN * 8)); }
          CK(cudaMemcpyAsync(h->d_notch_dumpblocks.p, blocks.data(), blocks.size() * 8, cudaMemcpyHostToDevice, h->st));
          fa.dump_blocks = h->d_notch_dumpblocks.as<uint64_t>(); fa.ndump = (int)blocks.size(); fa.dump = h->d_notch_dump.as<float2>();
          auto remap = [&](std::vector<uint64_t> &pts) {
            for (uint64_t &pnt : pts) {
              const size_t slot = std::lower_bound(blocks.begin(), blocks.end(), pnt / kNotchN) - blocks.begin();
              pnt = slot * kNotchN + pnt % kNotchN;
            }
          };
          remap(pts_cnr); remap(pts_spec);
        }
        h->meas_abs_next = abs0 + nblocks * kNotchN;
      }
    }
    fa.n = a;
    KL("notch_fir", launch_notch_fir(fa, nullptr, 0, h->d_notch_guess.as<float2>(), h->st));
  } else {
    KL("notch_apply", launch_notch_apply(a, nullptr, 0, h->d_notch_guess.as<float2>(), h->st));
  }
  // Verify entry(j) == exit(j-1) bit for bit.  Segments whose warm-up had not merged
  // with the true trajectory are re-run exactly from their predecessor's exit state,
  // all of them in one launch per round (a segment whose predecessor is also being
  // repaired waits for the next round).
  std::vector<float2> exitv;
  { int rcv = notch_verify_repair(h, a, exitv, v2 ? &fa : nullptr); if (rcv) return rcv; }
  for (int s = 0; s < c.anf; ++s) {
    h->notch.slot[s].est_re = exitv[(size_t)(a.nsegs - 1) * kNotchMaxSlots + s].x;
    h->notch.slot[s].est_im = exitv[(size_t)(a.nsegs - 1) * kNotchMaxSlots + s].y;
  }
  if (fused) {
    KL("fir_edges", launch_fir_edges(fa, h->st));
    h->s_notched.count = (uint64_t)h->fir_n;
    h->s_notched.fresh = 0;
    h->s_pp.count += ycount;
    h->s_pp.fresh = ycount;
    if (!pts_cnr.empty() || !pts_spec.empty()) {
      MeasSrc msrc;
      memset(&msrc, 0, sizeof msrc);
      msrc.rest.head = fa.dump; msrc.rest.head_count = (uint64_t)fa.ndump * kNotchN;
      msrc.fmt = 5; msrc.scale = 1.0f;
      int rcm;
      if (!pts_cnr.empty() && (rcm = meas_launch(h, h->m_cnr, msrc, pts_cnr))) return rcm;
      if (!pts_spec.empty() && (rcm = meas_launch(h, h->m_spec, msrc, pts_spec))) return rcm;
    }
    return LDVB_OK;
  }
  h->s_notched.count += nblocks * kNotchN;
  h->s_notched.fresh = nblocks * kNotchN;
  return LDVB_OK;
}

// --------------------------------------------------------------------- front end

// fir_filter retune from the demodulator's freq_tap (dsp.h:236-244), sampled once per batch (the reference
// samples it once per run() call).  *real_taps: every shifted tap has a zero imaginary part.
int fir_retune(ldvb_handle *h, int *real_taps) {
  if (h->use_fir) {
    const float new_freq = h->rx_state.freq_tap * h->fir_tap_mult;
    if (fabsf(h->fir_current_freq - new_freq) > h->fir_tol) {
      h->fir_shifted = shift_taps(h->fir_coeffs, new_freq);
      h->fir_current_freq = new_freq;
      CK(cudaMemcpyAsync(h->d_taps.p, h->fir_shifted.data(), h->fir_shifted.size() * 4, cudaMemcpyHostToDevice, h->st));
      CK(cudaStreamSynchronize(h->st));
    }
  }
  if (real_taps) {
    *real_taps = 1;
    for (int i = 0; i < h->fir_n; ++i) if (h->fir_shifted[2 * i + 1] != 0.0f) *real_taps = 0;
  }
  return LDVB_OK;
}

int run_frontend(ldvb_handle *h, const RawSrc &src, int fmt, uint64_t avail, uint64_t *consumed) {
  const ldvb_config &c = h->cfg;
  const uint32_t N = h->use_fir ? (uint32_t)h->fir_n : 0;
  const uint32_t D = (uint32_t)h->decim;
  uint64_t count;
  if (N) count = (avail >= N) ? (avail - N) / D : 0;  // dsp.h:246-247
  else count = avail / D;                              // generic.h:254
  *consumed = count * D;
  h->s_pp.fresh = 0;
  if (!count) return LDVB_OK;
  if (h->s_pp.count + count > h->s_pp.cap) return fail(h, LDVB_EOVERFLOW, "preprocessed stream overflow");
  { int rcr = fir_retune(h, nullptr); if (rcr) return rcr; }
  FrontendArgs a;
  memset(&a, 0, sizeof a);
  a.src = src; a.fmt = fmt; a.scale = c.float_scale;
  a.rot_lut = h->use_rot ? h->d_rot.as<float>() : nullptr;
  a.rot_index0 = h->rot_index;
  a.taps = h->d_taps.as<float2>();
  a.ntaps = N; a.decim = D;
  a.real_taps = 1;
  for (uint32_t i = 0; i < N; ++i) if (h->fir_shifted[2 * i + 1] != 0.0f) a.real_taps = 0;
  a.out = reinterpret_cast<float2 *>(h->s_pp.at(h->s_pp.count));
  a.count = count;
  KL("frontend", launch_frontend(a, h->st));
  h->rot_index = (uint32_t)((h->rot_index + *consumed) & 0xffffu);
  h->s_pp.count += count;
  h->s_pp.fresh = count;
  return LDVB_OK;
}

// ----------------------------------------------------------------------- receiver

// FAST receiver, part 1: cut the owned chunks into spans, run them, stitch the seams.
int rx_fast_launch(ldvb_handle *h, RxArgs &a, RxStitchArgs &sa, uint64_t nchunks_owned) {
  const ldvb_config &c = h->cfg;
  const uint64_t nchunks = nchunks_owned;
    // Span length: fill the machine (one lane per span, ~57 K resident lanes), at
    // least 4 chunks; 4 chunks of warm-up (timing and carrier loops re-converge
    // within ~200 symbols when freqw and the AGC are carried).
    uint32_t S = c.span_chunks;
    static const uint64_t target_env = [] { const char *e = getenv("LDVB_RX_SPANS"); return e ? strtoull(e, nullptr, 10) : 0ull; }();
    if (!h->rx_target_spans) h->rx_target_spans = rx_resident_lanes(h->rxp.slicer);   // one full wave of k_rx on this device
    const uint64_t target = target_env ? target_env : h->rx_target_spans;
    if (!S) S = (uint32_t)std::max<uint64_t>(4, (nchunks + target - 1) / target);
    const uint32_t W = c.warmup_chunks ? c.warmup_chunks : 4;
    a.span_chunks = S;
    a.warm_chunks = W;
    a.nspans = (uint32_t)((nchunks + S - 1) / S);
    a.span_cap = ((uint32_t)((S + kRxVerifyChunks + 1) * kRxChunk * h->rx_sym_per_sample) + 64 + 3u) & ~3u;   // 16-byte groups (k_rx: emit_word)
    if (a.nspans > h->rx_max_spans || (uint64_t)a.nspans * a.span_cap * 4 > h->d_rx_spans.bytes)
      return fail(h, LDVB_EOVERFLOW, "receiver span buffers too small for this batch");
    a.sym_out = h->d_rx_spans.as<uint32_t>();
    a.head_log = h->d_rx_head.as<RxSeamSym>();
    a.tail_log = h->d_rx_tail.as<RxSeamSym>();
    a.state_begin = h->d_rx_begin.as<RxState>();
    KL("rx", launch_rx(a, nullptr, 0, h->st));
    memset(&sa, 0, sizeof sa);
    sa.info = a.info; sa.head_log = a.head_log; sa.tail_log = a.tail_log;
    sa.nspans = a.nspans; sa.nrot = h->cst.nrotations; sa.nsymbols = h->cst.nsymbols;
    sa.rot_perm = h->d_rotperm.as<uint8_t>(); sa.omega = h->rxp.omega;
    sa.seams = h->d_rx_seams.as<RxSeam>();
    // Seam rule (include/leandvb_b200.h, "Receiver scheduling mode"): strict = every hard decision of the overlap
    // agrees and the loop states on both sides agree: phase within 1/16 of the ambiguity sector (5.6 degrees for
    // QPSK; converged loops differ by a few phase units), NCO frequency within 1/64 of the lock range.
    sa.strict = (c.seam_mode == 0) ? 1 : 0;
    sa.state_begin = a.state_begin; sa.state_end = a.state_end;
    sa.tol_phase = c.hs ? 0.f : 65536.0f / (float)h->cst.nrotations / 16.0f;
    sa.tol_freqw = c.hs ? 0.f : (h->rxp.max_freqw - h->rxp.min_freqw) / 64.0f;
    KL("rx_stitch", launch_rx_stitch(sa, nullptr, 0, h->st));
  // The device is busy with the spans for a while: finish the telemetry of this batch now.
  { int rcm = meas_finish(h); if (rcm) return rcm; }
  return LDVB_OK;
}

// Seam verdicts and span counts of the last rx / rx_stitch launches.
int rx_fetch(ldvb_handle *h, const RxArgs &a, const RxStitchArgs &sa, std::vector<RxSeam> &seams,
             std::vector<RxSpanInfo> &info) {
  seams.resize(a.nspans); info.resize(a.nspans);
  if (a.nspans > 1) CK(cudaMemcpyAsync(seams.data(), sa.seams, sizeof(RxSeam) * (a.nspans - 1), cudaMemcpyDeviceToHost, h->st));
  CK(cudaMemcpyAsync(info.data(), a.info, sizeof(RxSpanInfo) * a.nspans, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return LDVB_OK;
}

// Cold start with a carrier offset: the carried freqw (0) is far from the truth, most
// warm-ups do not converge and most seams fail.  Re-seed the warm-up state with the
// median frequency / power that the spans themselves reached and run them again; the
// loops pull in a little more each pass (span 0 keeps the exact carried state when it
// has one).
int rx_fast_reseed(ldvb_handle *h, RxArgs &a, RxStitchArgs &sa, std::vector<RxSeam> &seams,
                   std::vector<RxSpanInfo> &info) {
  for (int attempt = 0; attempt < 6 && a.nspans > 8; ++attempt) {
    uint32_t nfail = 0;
    for (uint32_t j = 0; j + 1 < a.nspans; ++j) nfail += seams[j].ok_loose ? 0 : 1;   // spans that did not converge
    if (nfail <= std::max<uint32_t>(4, a.nspans / 32)) break;
    std::vector<RxState> ends(a.nspans);
    CK(cudaMemcpyAsync(ends.data(), a.state_end, sizeof(RxState) * a.nspans, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    std::vector<float> fw(a.nspans), pw(a.nspans);
    for (uint32_t j = 0; j < a.nspans; ++j) { fw[j] = ends[j].freqw; pw[j] = ends[j].est_insp; }
    std::nth_element(fw.begin(), fw.begin() + fw.size() / 2, fw.end());
    std::nth_element(pw.begin(), pw.begin() + pw.size() / 2, pw.end());
    RxState warm = h->rx_state;
    warm.freqw = fw[fw.size() / 2];
    warm.est_insp = pw[pw.size() / 2];
    if (warm.est_insp > 0) warm.agc_gain = 75.0f / sqrtf(warm.est_insp);
    CK(cudaMemcpyAsync(h->d_rx_forced.p, &warm, sizeof warm, cudaMemcpyHostToDevice, h->st));
    a.warm_in = h->d_rx_forced.as<RxState>();
    ++h->meas.seams_repaired;   // counted as one (global) repair pass
    KL("rx", launch_rx(a, nullptr, 0, h->st));
    KL("rx_stitch", launch_rx_stitch(sa, nullptr, 0, h->st));
    int rc = rx_fetch(h, a, sa, seams, info);
    if (rc) return rc;
  }
  return LDVB_OK;
}

// FAST receiver, part 2: resolve the seams (device plan, host repair on failure), write the
// contiguous symbol stream at sym_dst and carry the loop state.  rot0 / skip0 describe the
// seam in front of span 0 (time-sharded mode; 0 otherwise).
int rx_fast_resolve(ldvb_handle *h, RxArgs &a, RxStitchArgs &sa, int rot0, uint32_t skip0, uint32_t *sym_dst,
                    uint64_t room, uint64_t *produced_out, int *cum_out, const std::function<int()> *pre_sync = nullptr) {
  uint64_t produced = 0;
  if (!h->rx_mail && cudaHostAlloc((void **)&h->rx_mail, sizeof(RxMail), cudaHostAllocDefault) != cudaSuccess)
    return fail(h, LDVB_ENOMEM, "receiver mailbox");
  RxMail *mail = h->rx_mail;
    // Fast path: offsets / skips / rotations are resolved on the device; the host reads back
    // four numbers.  Only when a seam failed (or a span overflowed) the seams are fetched
    // and repaired below.
    KL("rx_plan", launch_rx_plan(a.info, sa.seams, a.nspans, a.span_cap, h->cst.nrotations, rot0, skip0,
                                 h->d_rx_off.as<uint64_t>(), h->d_rx_skip.as<uint32_t>(), h->d_rx_rot.as<uint8_t>(),
                                 h->d_counts.as<uint64_t>(), h->st));
    // One synchronisation for the plan, the end state of the last span (final unless a seam has to be repaired) and
    // whatever else the caller wants to read back with them.
    uint64_t *plan = mail->plan;
    CK(cudaMemcpyAsync(plan, h->d_counts.p, 9 * 8, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(&mail->last, a.state_end + (a.nspans - 1), sizeof(RxState), cudaMemcpyDeviceToHost, h->st));
    if (pre_sync) { int rcp = (*pre_sync)(); if (rcp) return rcp; }
    CK(cudaStreamSynchronize(h->st));
    int cum = (int)plan[2];
    const bool slow = (plan[0] != 0 || plan[3] != 0);
    h->meas.seams_total += slow ? 0 : a.nspans - 1;
    auto note_deviation = [&](const uint64_t *pl) {
      float f; uint32_t u;
      u = (uint32_t)pl[5]; memcpy(&f, &u, 4); h->meas.seam_max_dphase = std::max(h->meas.seam_max_dphase, f);
      u = (uint32_t)pl[6]; memcpy(&f, &u, 4); h->meas.seam_max_dfreqw = std::max(h->meas.seam_max_dfreqw, f);
      u = (uint32_t)pl[7]; memcpy(&f, &u, 4); h->meas.seam_max_dmu = std::max(h->meas.seam_max_dmu, f);
    };
    if (!slow) { h->meas.seams_mismatch_accepted += (uint32_t)plan[4]; note_deviation(plan); }
    produced = plan[1];
    if (slow) {
    std::vector<RxSeam> seams(a.nspans);
    std::vector<RxSpanInfo> info(a.nspans);
    auto fetch = [&]() -> int { return rx_fetch(h, a, sa, seams, info); };
    int rc = fetch();
    if (rc) return rc;
    if ((rc = rx_fast_reseed(h, a, sa, seams, info))) return rc;
    h->meas.seams_total += a.nspans - 1;
    // Repair failed seams: span j+1 is re-run exactly from the end state of span j.  All
    // failed spans whose predecessor is final are repaired in ONE launch per round; the
    // seam behind each repaired span is stitched again (its tail changed).
    // Strict mode: after kStrictRounds rounds the seams that still carry isolated mismatches are judged
    // by the tolerant rule (low SNR: two converged loops flip different noise-level decisions, and every
    // round costs a span time); seams that fail that rule too are repaired to the end.
    constexpr int kStrictRounds = 2;
    std::vector<uint8_t> fixed(a.nspans, 0);   // seam j resolved by an exact re-run of span j+1
    for (int round = 0; round < 1 << 20; ++round) {
      std::vector<uint32_t> spans, restitch;
      bool prev_bad = false;
      const bool loose = sa.strict && round >= kStrictRounds;
      for (uint32_t j = 0; j + 1 < a.nspans; ++j) {
        const bool bad = !(loose ? seams[j].ok_loose : seams[j].ok) && !fixed[j];
        if (bad && !prev_bad) spans.push_back(j + 1);
        prev_bad = bad;
      }
      if (spans.empty()) break;
      h->meas.seams_repaired += (uint32_t)spans.size();
      for (uint32_t sp : spans) {
        fixed[sp - 1] = 1;
        if (sp + 1 < a.nspans) { fixed[sp] = 0; restitch.push_back(sp); }
      }
      if (h->d_scratch.bytes < (spans.size() + restitch.size()) * 4) return fail(h, LDVB_EOVERFLOW, "repair list");
      uint32_t *d_list = h->d_scratch.as<uint32_t>();
      CK(cudaMemcpyAsync(d_list, spans.data(), spans.size() * 4, cudaMemcpyHostToDevice, h->st));
      KL("rx", launch_rx(a, d_list, (uint32_t)spans.size(), h->st));
      if (!restitch.empty()) {
        CK(cudaMemcpyAsync(d_list + spans.size(), restitch.data(), restitch.size() * 4, cudaMemcpyHostToDevice, h->st));
        KL("rx_stitch", launch_rx_stitch(sa, d_list + spans.size(), (uint32_t)restitch.size(), h->st));
      }
      rc = fetch();
      if (rc) return rc;
    }
    for (uint32_t j = 0; j + 1 < a.nspans; ++j) {
      if (!fixed[j]) {
        if (seams[j].mismatches) ++h->meas.seams_mismatch_accepted;
        h->meas.seam_max_dphase = std::max(h->meas.seam_max_dphase, fabsf(seams[j].dphase));
        h->meas.seam_max_dfreqw = std::max(h->meas.seam_max_dfreqw, fabsf(seams[j].dfreqw));
        h->meas.seam_max_dmu = std::max(h->meas.seam_max_dmu, fabsf(seams[j].dmu));
      }
    }
    for (uint32_t j = 0; j + 1 < a.nspans; ++j)
      if (fixed[j]) { seams[j].ok = 1; seams[j].rot = 0; seams[j].extend_prev = 0; seams[j].skip_next = 0; }
    // Offsets, skips and cumulative rotations.
    std::vector<uint64_t> off(a.nspans + 1, 0);
    std::vector<uint32_t> skipv(a.nspans, 0);
    std::vector<uint8_t> rot(a.nspans, 0);
    cum = rot0 % h->cst.nrotations;
    skipv[0] = skip0;
    for (uint32_t j = 0; j < a.nspans; ++j) {
      if (info[j].n_out + info[j].n_tail > a.span_cap) return fail(h, LDVB_EOVERFLOW, "span capacity exceeded");
      uint64_t keep = info[j].n_out;
      if (j > 0) {
        skipv[j] = (uint32_t)seams[j - 1].skip_next;
        cum = (cum + seams[j - 1].rot) % h->cst.nrotations;
      }
      rot[j] = (uint8_t)cum;
      keep -= skipv[j];
      if (j + 1 < a.nspans) keep += (uint64_t)seams[j].extend_prev;
      off[j + 1] = off[j] + keep;
    }
    produced = off[a.nspans];
    if (produced > room) return fail(h, LDVB_EOVERFLOW, "symbol stream overflow");
    CK(cudaMemcpyAsync(h->d_rx_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->d_rx_skip.p, skipv.data(), skipv.size() * 4, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->d_rx_rot.p, rot.data(), rot.size(), cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));   // the vectors above are locals
    }  // slow path
    if (produced > room) return fail(h, LDVB_EOVERFLOW, "symbol stream overflow");
    RxCompactArgs ca;
    ca.sym_in = a.sym_out; ca.span_cap = a.span_cap; ca.nspans = a.nspans;
    ca.span_offset = h->d_rx_off.as<uint64_t>(); ca.span_skip = h->d_rx_skip.as<uint32_t>();
    ca.span_rot = h->d_rx_rot.as<uint8_t>(); ca.rot_perm = h->d_rotperm.as<uint8_t>();
    ca.nsymbols = h->cst.nsymbols; ca.sym_out = sym_dst;
    KL("rx_compact", launch_rx_compact(ca, produced, h->st));
    // Carry: the end state of the last span.  Its phase is rotated back by the
    // cumulative rotation so that the next batch continues in span 0's frame.
    if (slow) {      // spans were re-run: fetch the state again
      CK(cudaMemcpyAsync(&mail->last, a.state_end + (a.nspans - 1), sizeof(RxState), cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
    }
    h->rx_state = mail->last;
    if (cum) {
      // Span frames differ by cum*65536/nrot phase units: symbols of the last span
      // were de-rotated by `cum`; adding the same angle to the PLL phase makes the
      // next batch's span 0 produce symbols in the reference frame directly.
      float shift = (float)cum * (65536.0f / h->cst.nrotations);
      h->rx_state.phase = fmodf(h->rx_state.phase - shift, 65536.0f);
    }
    *produced_out = produced;
  *cum_out = cum;
  return LDVB_OK;
}

// Level check in front of the FAST spans (include/leandvb_b200.h, "AGC settling"): mean |x|^2 of the first samples
// of the batch against the carried AGC estimate (sdr.h:863-869 tracks the power of the sampled symbols).
// *far = the estimate is more than a factor 2 away, i.e. the spans' warm-ups would start with loop gains
// (proportional to the squared amplitude) that are off by more than that.
int rx_level_check(ldvb_handle *h, const float2 *x, uint64_t nsamples, float *power, bool *far) {
  *far = false; *power = 0;
  if (h->cfg.hs || h->cfg.settle_chunks < 0 || !nsamples) return LDVB_OK;   // --hs: no AGC in that receiver
  KL("rx_power", launch_rx_power(x, (uint32_t)std::min<uint64_t>(nsamples, 16384), h->d_rx_power.as<float>(), h->st));
  CK(cudaMemcpyAsync(power, h->d_rx_power.p, 4, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  const float est = h->rx_state.est_insp;
  *far = (*power > 0.f) && (est > 2.0f * *power || est < 0.5f * *power);
  return LDVB_OK;
}

uint64_t rx_settle_chunks(const ldvb_handle *h) { return h->cfg.settle_chunks > 0 ? (uint64_t)h->cfg.settle_chunks : 512; }

// The settling pass: chunks [a.chunk0, a.chunk0 + K) are walked by the serial lane from the exact carried state
// (a.state_in), symbols go straight to sym_dst.  On return `a` describes the rest of the batch: it starts at
// chunk0 + K from the state reached there (kept on the device), exactly, like span 0 of an ordinary batch.
int rx_settle(ldvb_handle *h, RxArgs &a, uint64_t K, uint32_t *sym_dst, uint64_t room, uint64_t *n0) {
  RxArgs s = a;
  RxState *d_state = h->d_rx_settled.as<RxState>();
  RxSpanInfo *d_info = reinterpret_cast<RxSpanInfo *>(h->d_rx_settled.as<uint8_t>() + ((sizeof(RxState) + 15) & ~(size_t)15));
  s.nchunks = a.chunk0 + K; s.avail_chunks = s.nchunks;
  s.span_chunks = (uint32_t)K; s.warm_chunks = 0; s.nspans = 1;
  s.span_cap = (uint32_t)std::min<uint64_t>(room, 0xffffffffu);
  s.sym_out = sym_dst; s.info = d_info; s.state_end = d_state;
  s.head_log = nullptr; s.tail_log = nullptr; s.state_begin = nullptr; s.sampled = nullptr; s.sampled_flag = nullptr;
  KL("rx_settle", launch_rx(s, nullptr, 0, h->st));
  RxSpanInfo inf;
  CK(cudaMemcpyAsync(&inf, d_info, sizeof inf, cudaMemcpyDeviceToHost, h->st));
  CK(cudaMemcpyAsync(&h->rx_state, d_state, sizeof(RxState), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (inf.n_out > s.span_cap) return fail(h, LDVB_EOVERFLOW, "symbol stream overflow");
  *n0 = inf.n_out;
  ++h->meas.settle_passes;
  a.chunk0 += K;
  a.state_in = d_state; a.warm_in = d_state; a.state_chunk = a.chunk0;
  a.first_exact = 1;
  return LDVB_OK;
}

int run_receiver(ldvb_handle *h) {
  const ldvb_config &c = h->cfg;
  Stream &in = h->s_pp;
  h->s_sym.fresh = 0;
  h->meas_log.clear();
  if (in.count < (uint64_t)kRxChunk + h->readahead) return LDVB_OK;
  const uint64_t nchunks = (in.count - h->readahead) / kRxChunk;  // sdr.h:783
  RxArgs a;
  memset(&a, 0, sizeof a);
  a.p = h->rxp;
  a.p.cstln = h->d_cstln.as<CstlnCellDev>();
  a.p.trig = h->d_trig.as<float2>();
  a.x = reinterpret_cast<const float2 *>(in.at(0));
  a.nchunks = nchunks; a.avail_chunks = nchunks;
  a.chunk0 = 0; a.first_exact = 1; a.prev_end = nullptr;
  CK(cudaMemcpyAsync(h->d_rx_state.p, &h->rx_state, sizeof(RxState), cudaMemcpyHostToDevice, h->st));
  CK(cudaMemsetAsync(h->d_rx_measn.p, 0, 4, h->st));
  a.state_in = h->d_rx_state.as<RxState>();
  a.warm_in = a.state_in;
  a.info = h->d_rx_info.as<RxSpanInfo>();
  a.state_end = h->d_rx_end.as<RxState>();
  a.meas = h->d_rx_meas.as<float>();
  a.meas_count = h->d_rx_measn.as<uint32_t>();
  a.max_meas = 4096;
  uint32_t *sym_dst = reinterpret_cast<uint32_t *>(h->s_sym.at(h->s_sym.count));
  const uint64_t room = h->s_sym.cap - h->s_sym.count;
  uint64_t produced = 0;
  const bool fast = (c.rx_mode == LDVB_RX_FAST) && nchunks >= 4;
  bool fast_rows_done = false;
  if (!fast) {
    a.span_chunks = (uint32_t)std::min<uint64_t>(nchunks, 0xffffffffu);
    a.warm_chunks = 0;
    a.nspans = 1;
    a.span_cap = (uint32_t)std::min<uint64_t>(room, 0xffffffffu);
    a.sym_out = sym_dst;
    DevBuf smp, smpf;
    if (c.keep_taps) {
      CK(smp.alloc(nchunks * 8)); CK(smpf.alloc(nchunks * 4));
      a.sampled = smp.as<float2>(); a.sampled_flag = smpf.as<uint32_t>();
    }
    KL("rx", launch_rx(a, nullptr, 0, h->st));
    RxSpanInfo inf;
    CK(cudaMemcpyAsync(&inf, a.info, sizeof inf, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(&h->rx_state, a.state_end, sizeof(RxState), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (inf.n_out > a.span_cap) { smp.release(); smpf.release(); return fail(h, LDVB_EOVERFLOW, "symbol stream overflow"); }
    produced = inf.n_out;
    if (c.keep_taps) {
      // p_sampled holds one entry per chunk that produced a symbol (sdr.h:857-861)
      std::vector<float2> sv(nchunks); std::vector<uint32_t> fv(nchunks);
      CK(cudaMemcpy(sv.data(), smp.p, nchunks * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(fv.data(), smpf.p, nchunks * 4, cudaMemcpyDeviceToHost));
      std::vector<float2> packed;
      for (uint64_t i = 0; i < nchunks; ++i) if (fv[i]) packed.push_back(sv[i]);
      Tap &t = h->taps[LDVB_TAP_SAMPLED];
      if (t.buf.bytes < packed.size() * 8) { t.buf.release(); CK(t.buf.alloc(packed.size() * 8 + 4096)); }
      CK(cudaMemcpy(t.buf.p, packed.data(), packed.size() * 8, cudaMemcpyHostToDevice));
      t.bytes = packed.size() * 8;
      smp.release(); smpf.release();
    }
  } else {
    int rcf;
    bool rows_fetched = false;
    // What is read back together with the seam plan (one synchronisation): the telemetry rows of the kernel and,
    // in the steady state, the level check.
    const std::function<int()> fetch_rows = [&]() -> int {
      RxMail *mail = h->rx_mail;
      CK(cudaMemcpyAsync(&mail->nm, h->d_rx_measn.p, 4, cudaMemcpyDeviceToHost, h->st));
      CK(cudaMemcpyAsync(mail->rows, h->d_rx_meas.p, sizeof mail->rows, cudaMemcpyDeviceToHost, h->st));
      if (h->d_rx_power.p) CK(cudaMemcpyAsync(&mail->power, h->d_rx_power.p, 4, cudaMemcpyDeviceToHost, h->st));
      rows_fetched = true;
      return LDVB_OK;
    };
    const float est_carried = h->rx_state.est_insp;      // (the estimate the spans of this batch start from)
    auto is_far = [&](float power) {
      return (power > 0.f) && (est_carried > 2.0f * power || est_carried < 0.5f * power);
    };
    const bool settle_on = !c.hs && c.settle_chunks >= 0;
    // AGC settling: an estimate far from the input level is pulled in by a serial (exact) pass first.
    // The first FAST batch of a stream always settles: spans restart from the carried AGC estimate and move it by
    // only ~9 % per batch (S + W chunks at k = 0.01), so an estimate that starts at the constructor's 75^2
    // (sdr.h:727) would stay 10-20 % off the serial one for many batches -- hard decisions do not care, soft
    // costs (proportional to the gain squared) do: measured 13 % mean cost deviation without this (bench.py,
    // fast_vs_exact.steady_state).
    // Steady state (the stream has settled): the level check does not gate the launch.  The power of the batch head
    // is measured on the device, the spans are launched from the carried state, and the verdict comes back with the
    // seam plan; in the rare case that the level has jumped by more than a factor 2 the batch is redone the slow way.
    bool redo = false;
    const RxState carried = h->rx_state;
    const ldvb_meas meas_before = h->meas;
    if (settle_on && !h->rx_cold && nchunks >= 16) {
      KL("rx_power", launch_rx_power(a.x, (uint32_t)std::min<uint64_t>(nchunks * kRxChunk, 16384), h->d_rx_power.as<float>(), h->st));
      RxArgs a2 = a;
      RxStitchArgs sa;
      if ((rcf = rx_fast_launch(h, a2, sa, nchunks))) return rcf;
      int cum = 0;
      uint64_t n1 = 0;
      if ((rcf = rx_fast_resolve(h, a2, sa, 0, 0, sym_dst, room, &n1, &cum, &fetch_rows))) return rcf;
      if (is_far(h->rx_mail->power)) {
        redo = true;                                   // nothing is committed yet: symbols are overwritten below
        h->rx_state = carried;
        h->meas = meas_before;
        rows_fetched = false;
        CK(cudaMemsetAsync(h->d_rx_measn.p, 0, 4, h->st));
      } else {
        produced = n1;
      }
    } else {
      redo = true;
    }
    if (redo) {
      float power = 0; bool far = false;
      if ((rcf = rx_level_check(h, a.x, nchunks * kRxChunk, &power, &far))) return rcf;
      uint64_t n0 = 0, K = 0;
      if (h->rx_cold && settle_on) far = true;
      if (far) {
        h->rx_cold = false;
        K = std::min(nchunks, rx_settle_chunks(h));
        if (nchunks - K < 8) K = nchunks;                  // too little left for spans
        if ((rcf = rx_settle(h, a, K, sym_dst, room, &n0))) return rcf;
      }
      produced = n0;
      if (K < nchunks) {
        RxStitchArgs sa;
        if ((rcf = rx_fast_launch(h, a, sa, nchunks - K))) return rcf;
        int cum = 0;
        uint64_t n1 = 0;
        if ((rcf = rx_fast_resolve(h, a, sa, 0, 0, sym_dst + n0, room - n0, &n1, &cum, &fetch_rows))) return rcf;
        produced += n1;
      }
    }
    if (rows_fetched) {
      const uint32_t nm = std::min(h->rx_mail->nm, 4096u);
      const float *m = h->rx_mail->rows;
      std::vector<std::pair<float, int>> order;
      for (uint32_t i = 0; i < nm; ++i) order.push_back({m[4 * i], (int)i});
      std::sort(order.begin(), order.end());
      for (auto &o : order) {
        h->meas_log.push_back(m[4 * o.second + 1]);
        h->meas_log.push_back(m[4 * o.second + 2]);
        const float q = m[4 * o.second + 3];                               // est_sp / est_ep, -1: est_ep == 0
        h->meas_log.push_back(q < 0 ? 0.0f : 10 * logf(q) / logf(10));     // sdr.h:910-911
      }
    }
    fast_rows_done = rows_fetched;
  }
  // Measurements recorded by the kernel: {chunk, freq_tap, ss, mer}
  if (!fast_rows_done) {
    uint32_t nm = 0;
    CK(cudaMemcpy(&nm, h->d_rx_measn.p, 4, cudaMemcpyDeviceToHost));
    nm = std::min(nm, 4096u);
    if (nm) {
      std::vector<float> m(4 * nm);
      CK(cudaMemcpy(m.data(), h->d_rx_meas.p, m.size() * 4, cudaMemcpyDeviceToHost));
      std::vector<std::pair<float, int>> order;
      for (uint32_t i = 0; i < nm; ++i) order.push_back({m[4 * i], (int)i});
      std::sort(order.begin(), order.end());
      for (auto &o : order) {
        h->meas_log.push_back(m[4 * o.second + 1]);
        h->meas_log.push_back(m[4 * o.second + 2]);
        const float q = m[4 * o.second + 3];                               // est_sp / est_ep, -1: est_ep == 0
        h->meas_log.push_back(q < 0 ? 0.0f : 10 * logf(q) / logf(10));     // sdr.h:910-911

      }
    }
  }
  h->s_sym.count += produced;
  h->s_sym.fresh = produced;
  h->meas.symbols += produced;
  int rc = stream_consume(h, in, nchunks * kRxChunk, h->d_scratch);
  if (rc) return rc;
  h->meas.freq_tap = h->rx_state.freq_tap;
  h->meas.ss = sqrtf(h->rx_state.est_insp);
  h->meas.mer = h->rx_state.est_ep ? 10 * logf(h->rx_state.est_sp / h->rx_state.est_ep) / logf(10) : 0;
  return LDVB_OK;
}

// --------------------------------------------------------- deconvolution + sync

// One run of the deconvolver of the locked hypothesis over the unread symbols.
// Bytes are written behind the current end of s_bytes but nothing is committed:
// the caller appends them, lets the sync tracker look at them and only then
// commits the shift-register state and drops the symbols (deconv_commit).
struct DeconvRun { uint64_t produced = 0, consumed = 0; HypState after; };

int deconv_launch(ldvb_handle *h, uint64_t limit_bytes, DeconvRun *run) {
  Stream &in = h->s_sym;
  *run = DeconvRun();
  run->after = h->hyp[h->locked];
  if (in.count < 64) return LDVB_OK;  // dvb.h:419
  const int pp = h->dec.punctperiod, pw = h->dec.punctweight;
  uint64_t n = (in.count - 64) / (pw / 2) * pp / 8;  // dvb.h:420
  n = std::min(n, limit_bytes);
  n = std::min<uint64_t>(n, h->s_bytes.cap - h->s_bytes.count);
  if (!n) return LDVB_OK;
  const HypState &hs = h->hyp[h->locked];
  DeconvArgs a;
  memset(&a, 0, sizeof a);
  a.symbols = reinterpret_cast<const uint32_t *>(in.at(0));
  a.nbytes = n;
  a.reg_in = hs.reg; a.n_in = hs.n_in; a.out_acc = hs.acc; a.n_out = hs.n_out;
  for (int s = 0; s < 4; ++s) a.hyp[s] = h->dec.hyp_lut[h->locked][s];
  a.punctperiod = pp; a.punctweight = pw;
  for (int b = 0; b < 8; ++b) a.deconv[b] = h->dec.deconv[b];
  a.out = h->s_bytes.at(h->s_bytes.count);
  KL("deconv_carry", launch_deconv_carry(a, in.count, h->d_deconv_carry.as<uint64_t>(), h->st));
  uint64_t carry[5];
  CK(cudaMemcpyAsync(carry, h->d_deconv_carry.p, sizeof carry, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  run->after.reg = carry[0]; run->after.n_in = (int)(int64_t)carry[1];
  run->after.acc = carry[2]; run->after.n_out = (int)(int64_t)carry[3];
  run->consumed = carry[4];
  run->produced = n;
  if (run->consumed > in.count) return fail(h, LDVB_ESTATE, "deconvolver over-consumed");
  return LDVB_OK;
}

// --fastlock (dvb.h:428-454): one reference run() sees a window of n bytes, counts for each of the
// four alignments the disagreements between the deconvolution polynomials and their alternates on
// that alignment's AUXILIARY register (readerrors, dvb.h:391-412), locks the best one and asks for a
// one-symbol skip when even the best one is wrong on more than a third of the bits.  Here a window is
// what one batch lets the deconvolver produce; while the alignment is still wrong it is cut to
// kFastlockProbe bytes so that the skip takes effect after a window of about the size the reference
// sees with its default buffers instead of after a whole batch.
constexpr uint64_t kFastlockProbe = 1024;

uint64_t deconv_window(ldvb_handle *h) {
  const Stream &in = h->s_sym;
  if (in.count < 64) return 0;                                                        // dvb.h:419
  uint64_t n = (in.count - 64) / (h->dec.punctweight / 2) * h->dec.punctperiod / 8;  // dvb.h:420
  return std::min<uint64_t>(n, h->s_bytes.cap - h->s_bytes.count);
}

int fastlock_evaluate(ldvb_handle *h, uint64_t n, uint64_t *errors_best) {
  Stream &in = h->s_sym;
  uint64_t *dev = h->d_deconv_carry.as<uint64_t>();   // [4] error counts, [4][5] carries
  CK(cudaMemsetAsync(dev, 0, 4 * 8, h->st));
  for (int k = 0; k < 4; ++k) {
    DeconvArgs a;
    memset(&a, 0, sizeof a);
    a.symbols = reinterpret_cast<const uint32_t *>(in.at(0));
    a.nbytes = n;
    a.reg_in = h->hyp2[k].reg; a.n_in = h->hyp2[k].n_in; a.out_acc = 0; a.n_out = h->hyp2[k].n_out;
    for (int s = 0; s < 4; ++s) a.hyp[s] = h->dec.hyp_lut[k][s];
    a.punctperiod = h->dec.punctperiod; a.punctweight = h->dec.punctweight;
    for (int b = 0; b < 8; ++b) { a.deconv[b] = h->dec.deconv[b]; a.deconv2[b] = h->dec.deconv2[b]; }
    a.err_out = reinterpret_cast<unsigned long long *>(dev + k);
    KL("deconv_errors", launch_deconv_carry(a, in.count, dev + 4 + 5 * k, h->st));
  }
  uint64_t host[24];
  CK(cudaMemcpyAsync(host, dev, sizeof host, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  uint64_t best_err = 1ull << 30; int best = 0;
  for (int k = 0; k < 4; ++k) {
    if (host[k] < best_err) { best_err = host[k]; best = k; }                          // dvb.h:436-439
    const uint64_t *c = host + 4 + 5 * k;
    h->hyp2[k].reg = c[0]; h->hyp2[k].n_in = (int)(int64_t)c[1]; h->hyp2[k].n_out = (int)(int64_t)c[3];
  }
  h->locked = best;                                                                    // dvb.h:441-447
  *errors_best = best_err;
  return LDVB_OK;
}

int deconv_commit(ldvb_handle *h, const DeconvRun &run) {
  h->hyp[h->locked] = run.after;
  return stream_consume_lazy(h, h->s_sym, run.consumed);
}

// viterbi_sync::run (dvb.h:1366-1414): whole chunks of 128 FEC blocks.
int run_viterbi(ldvb_handle *h, uint64_t *produced) {
  Stream &in = h->s_sym;
  *produced = 0;
  const int nsh = h->vsyncs.nshifts, bits_in = h->trellis.bits_in;
  const uint64_t need = (uint64_t)nsh * 128 + (nsh - 1);
  if (in.count < need) return LDVB_OK;
  uint64_t nchunks = (in.count - (nsh - 1)) / ((uint64_t)nsh * 128);
  const uint64_t bpc = (uint64_t)16 * bits_in;   // bytes per chunk
  nchunks = std::min(nchunks, (h->s_bytes.cap - h->s_bytes.count) / bpc);
  if (!nchunks) return LDVB_OK;
  VitArgs a;
  memset(&a, 0, sizeof a);
  a.symbols = reinterpret_cast<const uint32_t *>(in.at(0));
  a.nchunks = nchunks;
  a.bits_in = bits_in; a.bits_out = h->trellis.bits_out; a.bps = h->vsyncs.bps; a.nshifts = nsh;
  a.nsyncs = h->vsyncs.nsyncs; a.ncs = h->trellis.ncs; a.nsymbols = h->cst.nsymbols;
  a.path_nbits = h->trellis.path_nbits; a.path_depth = h->trellis.path_depth; a.path32 = h->trellis.path32 ? 1 : 0;
  a.resync_period = h->cfg.fastlock ? 1 : 32;   // dvb.h:1241, leandvb.cc:540
  a.trellis_pred = h->d_vit_pred.as<uint8_t>(); a.trellis_us = h->d_vit_us.as<uint8_t>();
  a.maps = h->d_vit_maps.as<uint8_t>(); a.shifts = h->d_vit_shifts.as<int32_t>();
  a.state = h->d_vit_state.as<VitDecState>(); a.ctl = h->d_vit_ctl.as<VitCtl>();
  a.out = h->s_bytes.at(h->s_bytes.count);
  {
    // Time segments (k_viterbi.cu): whole re-sync groups, cold start + warm-up, verified bit for bit.
    const uint64_t P = (uint64_t)a.resync_period;
    const bool no_warm = h->cfg.vit_warm_chunks < 0;
    // Warm-up: 2 chunks bring the current decoder (right hypothesis) back; the decoders of wrong hypotheses
    // decode noise and need ~2000 blocks = 16 chunks (they only run on re-sync chunks: 16 of those; with 8,
    // 11-13 of ~1100-2000 segments still failed, and a repair round costs a whole segment time).
    const uint32_t warm_others = no_warm ? 0 : 16;
    const uint32_t warm = no_warm ? 0 : h->cfg.vit_warm_chunks ? (uint32_t)h->cfg.vit_warm_chunks : (P > 1 ? 2 : 16);
    VitCtl ctl0;
    CK(cudaMemcpyAsync(&ctl0, h->d_vit_ctl.p, sizeof ctl0, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    // Default: one full wave of CTAs (every segment resident at once, no tail).
    if (!h->vit_wave) h->vit_wave = std::max(1, vit_resident_segments(a.ncs, bits_in, a.nsyncs));
    const uint64_t target = h->cfg.vit_segments > 0 ? (uint64_t)h->cfg.vit_segments : (uint64_t)h->vit_wave;
    const uint64_t minL = std::max<uint64_t>(P, 128);
    uint64_t L = std::max<uint64_t>(minL, (nchunks + target - 1) / target);
    L = (L + P - 1) / P * P;
    std::vector<uint64_t> start;
    start.push_back(0);
    // first boundary: a re-sync chunk, at least one group (and the warm-up) into the batch
    uint64_t b = (P - (uint64_t)ctl0.resync_phase % P) % P;
    // (P > 1: a segment closer than warm_others re-sync chunks to the start takes the other decoders exactly
    //  from the carried state, so only the current decoder's warm-up has to fit in front of it)
    while (b < std::max<uint64_t>(P > 1 ? P : 1, warm) || b < L / 2) b += P;
    if (target > 1) for (; b + L / 2 < nchunks; b += L) start.push_back(b);
    start.push_back(nchunks);
    const uint32_t nseg = (uint32_t)start.size() - 1;
    const size_t st_bytes = (size_t)nseg * a.nsyncs * sizeof(VitDecState);
    if (h->d_vit_entry.bytes < st_bytes) {
      h->d_vit_entry.release(); h->d_vit_exit.release();
      CK(h->d_vit_entry.alloc(st_bytes + 4096)); CK(h->d_vit_exit.alloc(st_bytes + 4096));
    }
    const size_t aux = (size_t)nseg * (2 * sizeof(VitCtl) + 1 + 4) + (start.size()) * 8 + 256;
    if (h->d_vit_aux.bytes < aux) { h->d_vit_aux.release(); CK(h->d_vit_aux.alloc(aux + 4096)); }
    uint8_t *ax = h->d_vit_aux.as<uint8_t>();
    VitSegArgs sg;
    memset(&sg, 0, sizeof sg);
    sg.seg_start = reinterpret_cast<const uint64_t *>(ax); ax += start.size() * 8;
    sg.ctl_entry = reinterpret_cast<VitCtl *>(ax); ax += (size_t)nseg * sizeof(VitCtl);
    sg.ctl_exit = reinterpret_cast<VitCtl *>(ax); ax += (size_t)nseg * sizeof(VitCtl);
    uint32_t *d_list = reinterpret_cast<uint32_t *>(ax); ax += (size_t)nseg * 4;
    uint32_t *d_nfail = reinterpret_cast<uint32_t *>(ax); ax += 8;
    uint8_t *d_ok = ax;
    sg.nseg = nseg; sg.list = nullptr; sg.nlist = 0; sg.warm_chunks = warm; sg.warm_others = P > 1 ? warm_others : 0;
    sg.phase0 = ctl0.resync_phase; sg.nb = vit_rescan_entries(bits_in);
    sg.full = vit_trellis_is_full(h->trellis.pred.data(), a.ncs, bits_in) ? 1 : 0;
    sg.entry = h->d_vit_entry.as<VitDecState>(); sg.exit = h->d_vit_exit.as<VitDecState>();
    CK(cudaMemcpyAsync(const_cast<uint64_t *>(sg.seg_start), start.data(), start.size() * 8, cudaMemcpyHostToDevice, h->st));
    KL("viterbi", launch_viterbi(a, sg, nseg, h->st));
    h->meas.vit_segments += nseg;
    std::vector<uint8_t> ok(nseg);
    for (uint32_t round = 0; nseg > 1 && round <= nseg; ++round) {
      uint32_t nfail = 0;
      CK(cudaMemsetAsync(d_nfail, 0, 4, h->st));
      KL("vit_verify", launch_vit_verify(sg, a.nsyncs, d_ok, d_nfail, h->st));
      CK(cudaMemcpyAsync(&nfail, d_nfail, 4, cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
      if (!nfail) break;
      CK(cudaMemcpyAsync(ok.data(), d_ok, nseg, cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
      ok[0] = 1;
      std::vector<uint32_t> todo;
      for (uint32_t g = 1; g < nseg; ++g) if (!ok[g] && ok[g - 1]) todo.push_back(g);
      if (todo.empty()) return fail(h, LDVB_ESTATE, "viterbi segment repair made no progress");
      h->meas.vit_repaired += (uint32_t)todo.size();
      CK(cudaMemcpyAsync(d_list, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, h->st));
      VitSegArgs rp = sg;
      rp.list = d_list; rp.nlist = (uint32_t)todo.size();
      KL("viterbi", launch_viterbi(a, rp, rp.nlist, h->st));
      CK(cudaStreamSynchronize(h->st));      // `todo` is a local buffer
    }
    KL("vit_commit", launch_vit_commit(a, sg, h->st));
  }
  *produced = nchunks * bpc;
  h->s_bytes.count += *produced;
  h->s_bytes.fresh += *produced;
  return stream_consume(h, in, nchunks * 128 * (uint64_t)nsh, h->d_scratch);
}

int run_sync(ldvb_handle *h) {
  Stream &in = h->s_bytes;
  Stream &out = h->s_mpeg;
  for (int guard = 0; guard < 1000000; ++guard) {
    uint64_t npk = 0;
    CK(cudaMemcpyAsync(h->d_sync_state.p, &h->sync, sizeof(SyncState), cudaMemcpyHostToDevice, h->st));
    if (h->sync.synchronized) {
      npk = (in.count >= 205) ? (in.count - 1) / 204 : 0;  // dvb.h:843
      npk = std::min<uint64_t>(npk, (out.cap - out.count) / 204);
      if (!npk) break;
      KL("sync_flags", launch_sync_flags(in.at(0), npk, h->d_sync_state.as<SyncState>(), h->d_badwords.as<uint32_t>(), h->st));
    } else if (in.count < 204 * 8 + 1) {
      break;  // dvb.h:758
    }
    KL("sync_track", launch_sync_track(in.at(0), in.count, h->d_sync_state.as<SyncState>(),
                                       h->d_badwords.as<uint32_t>(), npk, h->d_sync_res.as<SyncResult>(), h->st));
    SyncResult r;
    CK(cudaMemcpyAsync(&r, h->d_sync_res.p, sizeof r, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (r.produced) {
      KL("realign", launch_realign(in.at(0), r.produced, h->sync.bitphase, h->sync.polarity, out.at(out.count), h->st));
      out.count += r.produced;
      out.fresh += r.produced;
    }
    h->sync = r.st;
    for (int e = 0; e < r.events && e < 16; ++e) h->meas.lock = r.event_val[e];
    int rc = stream_consume(h, in, r.consumed, h->d_scratch);
    if (rc) return rc;
    if (r.need_next_sync) {
      // mpeg_sync asks the deconvolver for its next hypothesis (dvb.h:771-778,
      // 185-193).  Bytes that the old hypothesis produced beyond the search
      // position are discarded: their symbols are still unread (see run_chain).
      return 1;
    }
    if (!r.consumed && !r.produced) break;
  }
  h->meas.locktime = h->sync.locktime;
  return LDVB_OK;
}

int run_fec(ldvb_handle *h, uint8_t *ts_dst, uint64_t ts_cap, uint64_t *ts_out) {
  Stream &in = h->s_mpeg;
  *ts_out = 0;
  uint64_t npk = (in.count >= 2448) ? (in.count - 2244) / 204 : 0;  // dvb.h:933
  npk = std::min(npk, h->ts_cap);
  if (!npk) return LDVB_OK;
  DeintRsArgs a;
  a.mpeg = in.at(0); a.npackets = npk;
  a.gf_exp = h->d_gfexp.as<uint8_t>(); a.gf_log = h->d_gflog.as<uint8_t>();
  a.rs_out = h->cfg.keep_taps ? h->d_rs204.as<uint8_t>() : nullptr;
  a.rts_out = h->d_rts.as<uint8_t>();
  a.flags = h->d_rsflags.as<int32_t>();
  KL("deint_rs", launch_deint_rs(a, h->st));
  DerandArgs d;
  d.rts = a.rts_out; d.npackets = npk; d.pattern = h->d_derand.as<uint8_t>();
  d.pos_in = h->derand_pos; d.ts_out = ts_dst; d.ts_cap = ts_cap;
  d.counts = h->d_counts.as<uint64_t>(); d.flags = a.flags; d.scratch = h->d_scratch.as<uint32_t>();
  int nl = 0;
  KL("derand", launch_derand(d, h->st, &nl));
  h->launches += (nl > 0 ? nl - 1 : 0);
  uint64_t counts[4];
  CK(cudaMemcpyAsync(counts, h->d_counts.p, sizeof counts, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (counts[0] > ts_cap) return fail(h, LDVB_EOVERFLOW, "TS output buffer too small");
  *ts_out = counts[0];
  h->derand_pos = (int)counts[2];
  h->meas.rs_bits += npk * 204 * 8;
  h->meas.rs_errs += counts[3];
  if (h->cfg.vber) {
    // rs_decoder::run (dvb.h:1004-1052) reports (nbits, nerrs); rate_estimator::run adds them and
    // emits num/den once den >= sample_size (generic.h:286-299) -- tested after every packet here.
    h->vber_flags.resize(npk * 2);
    CK(cudaMemcpyAsync(h->vber_flags.data(), a.flags, npk * 8, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    for (uint64_t p = 0; p < npk; ++p) {
      h->vber_num += h->vber_flags[2 * p + 1];
      h->vber_den += 204 * 8;
      if (h->vber_den >= h->vber_sample) {
        h->vber_queue.push_back((float)h->vber_num / h->vber_den);
        h->vber_num = h->vber_den = 0;
      }
    }
  }
  h->meas.ts_packets += counts[0];
  h->meas.ts_dropped += counts[1];
  int rc;
  if ((rc = tap_store(h, LDVB_TAP_RSPACKETS, h->d_rs204.p, npk * 204))) return rc;
  if ((rc = tap_store(h, LDVB_TAP_RTSPACKETS, h->d_rts.p, npk * 188))) return rc;
  if ((rc = tap_store(h, LDVB_TAP_RSFLAGS, h->d_rsflags.p, npk * 8))) return rc;
  return stream_consume(h, in, npk * 204, h->d_scratch);
}

// Deconvolution <-> sync -> de-interleave -> RS -> derandomise on the symbols in s_sym.
int run_backend(ldvb_handle *h, uint8_t *ts_dst, uint64_t ts_cap, uint64_t *ts_out) {
  const ldvb_config &c = h->cfg;
  int rc;
  WallTimer *wt = nullptr;
  struct WtGuard { WallTimer *&p; ~WtGuard() { delete p; } } wt_guard{wt};
  delete wt; wt = new WallTimer(h, "wall:deconv_sync");
  h->s_mpeg.fresh = 0;   // (run_sync may be called several times per batch: the tap covers all of them)
  // ---- deconvolution <-> sync (with the backward next_sync edge, dvb.h:771-778)
  std::vector<uint8_t> tap_bytes;  // bytes that stay in the stream, for the tap
  if (c.viterbi) {
    uint64_t produced = 0;
    const uint64_t at = h->s_bytes.count;
    if ((rc = run_viterbi(h, &produced))) return rc;
    if (c.keep_taps && produced) {
      tap_bytes.resize(produced);
      CK(cudaMemcpyAsync(tap_bytes.data(), h->s_bytes.at(at), produced, cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
    }
    // mpeg_sync has no deconvolver to poke in this mode (leandvb.cc:560: r_deconv == NULL)
    for (int guard = 0; guard < 1 << 20; ++guard) {
      rc = run_sync(h);
      if (rc < 0) return rc;
      if (rc == 0) break;
    }
  }
  if (c.hs) {
    // dvb_deconvol_sync_hard (dvb.h:633-660): whole 64-byte chunks, then mpeg_sync (fast search)
    Stream &in = h->s_sym;
    uint64_t nchunks = in.count / 512;
    nchunks = std::min<uint64_t>(nchunks, (h->s_bytes.cap - h->s_bytes.count) / 64);
    if (nchunks) {
      const int period = c.fastlock ? 1 : 32;                       // leandvb.cc:853
      const uint64_t ngroups = nchunks / period + 2;
      if (h->d_hs_errors.bytes < ngroups * 16) { h->d_hs_errors.release(); CK(h->d_hs_errors.alloc(ngroups * 16 + 4096)); }
      if (h->d_hs_lock.bytes < nchunks + 1) { h->d_hs_lock.release(); CK(h->d_hs_lock.alloc(nchunks + 4096)); }
      if (!h->d_hs_state.p) CK(h->d_hs_state.alloc(64));
      HsDeconvArgs a;
      a.symbols = reinterpret_cast<const uint32_t *>(in.at(0));
      a.nchunks = nchunks; a.hist = h->hs_hist; a.hist_valid = h->hs_hist_valid;
      a.resync_phase = h->hs_resync_phase; a.resync_period = period; a.locked = h->hs_locked;
      a.errors = h->d_hs_errors.as<uint32_t>(); a.lock_of_chunk = h->d_hs_lock.as<uint8_t>();
      a.out = h->s_bytes.at(h->s_bytes.count);
      a.state_out = h->d_hs_state.as<int32_t>();
      int nl = 0;
      KL("hs_deconv", launch_hs_deconv(a, h->st, &nl));
      h->launches += (nl > 0 ? nl - 1 : 0);
      // carry: lock, vote phase, the last 32 symbols of what was consumed
      int32_t locked = 0;
      uint32_t last[32];
      const uint64_t used = nchunks * 512;
      CK(cudaMemcpyAsync(&locked, h->d_hs_state.p, 4, cudaMemcpyDeviceToHost, h->st));
      CK(cudaMemcpyAsync(last, in.at(used - 32), sizeof last, cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
      h->hs_locked = locked;
      h->hs_resync_phase = (int)((h->hs_resync_phase + nchunks) % (uint64_t)period);
      uint64_t hist = 0;
      for (int i = 0; i < 32; ++i) hist = (hist << 2) | ((last[i] >> 16) & 3u);
      h->hs_hist = hist; h->hs_hist_valid = 32;
      const uint64_t produced = nchunks * 64;
      if (c.keep_taps) {
        tap_bytes.resize(produced);
        CK(cudaMemcpy(tap_bytes.data(), h->s_bytes.at(h->s_bytes.count), produced, cudaMemcpyDeviceToHost));
      }
      h->s_bytes.count += produced;
      h->s_bytes.fresh += produced;
      if ((rc = stream_consume(h, in, used, h->d_scratch))) return rc;
    }
    for (int guard = 0; guard < 1 << 20; ++guard) {
      rc = run_sync(h);
      if (rc < 0) return rc;
      if (rc == 0) break;
    }
  }
  // The loop ends when the symbols are drained below what the deconvolver needs (dvb.h:419-426); every pass
  // consumes symbols, a skipped symbol or at least moves to the next hypothesis, so the bound below (one pass per
  // 256 symbols + a constant) is never reached by a correct run: running into it is reported, not hidden.
  // While mpeg_sync is searching, one pass deconvolves only what a search can look at before it calls
  // next_sync() (3 sweeps x 8 bit phases x 1632 bytes, dvb.h:758-778) instead of the whole batch: on a stream that
  // never locks the cost is one pass over the symbols, not one pass per hypothesis switch.
  constexpr uint64_t kSearchWindow = 3 * 8 * 1632 + 4096;
  const uint64_t max_passes = c.viterbi || c.hs ? 0 : h->s_sym.count / 256 + 4096;
  uint64_t pass = 0;
  for (; pass < max_passes; ++pass) {
    if (h->skip) {  // dvb.h:415-416
      if (h->s_sym.count < (uint64_t)h->skip) break;
      if ((rc = stream_consume_lazy(h, h->s_sym, h->skip))) return rc;
      h->skip = 0;
    }
    DeconvRun run;
    if (c.fastlock) {
      uint64_t n = deconv_window(h);
      if (n < 32) break;                                                               // dvb.h:424-426
      HypState aux[4];
      for (int k = 0; k < 4; ++k) aux[k] = h->hyp2[k];
      const int locked_before = h->locked;
      uint64_t eb = 0;
      if ((rc = fastlock_evaluate(h, n, &eb))) return rc;
      if (eb > n * 8 / 3 && n > kFastlockProbe) {        // still misaligned: a short window, then look again
        for (int k = 0; k < 4; ++k) h->hyp2[k] = aux[k];
        h->locked = locked_before;
        n = kFastlockProbe;
        if ((rc = fastlock_evaluate(h, n, &eb))) return rc;
      }
      if (eb > n * 8 / 3) h->skip = 1;                                                 // dvb.h:449-453
      if ((rc = deconv_launch(h, n, &run))) return rc;
      const size_t tap_at0 = tap_bytes.size();
      if (c.keep_taps && run.produced) {
        tap_bytes.resize(tap_at0 + run.produced);
        CK(cudaMemcpy(tap_bytes.data() + tap_at0, h->s_bytes.at(h->s_bytes.count), run.produced, cudaMemcpyDeviceToHost));
      }
      h->s_bytes.count += run.produced;
      h->s_bytes.fresh += run.produced;
      if ((rc = deconv_commit(h, run))) return rc;
      for (int g2 = 0; g2 < 1 << 20; ++g2) {             // mpeg_sync never calls next_sync() here (dvb.h:751)
        rc = run_sync(h);
        if (rc < 0) return rc;
        if (rc == 0) break;
      }
      continue;                                          // more windows while symbols remain
    }
    const uint64_t limit = h->sync.synchronized ? ~0ull : kSearchWindow;
    if ((rc = deconv_launch(h, limit, &run))) return rc;
    if (!run.produced) break;                            // fewer symbols left than one run needs (dvb.h:419-426)
    const size_t tap_at = tap_bytes.size();
    if (c.keep_taps && run.produced) {
      tap_bytes.resize(tap_at + run.produced);
      CK(cudaMemcpy(tap_bytes.data() + tap_at, h->s_bytes.at(h->s_bytes.count), run.produced, cudaMemcpyDeviceToHost));
    }
    h->s_bytes.count += run.produced;
    h->s_bytes.fresh += run.produced;
    rc = run_sync(h);
    if (rc < 0) return rc;
    if (rc == 0) {
      if ((rc = deconv_commit(h, run))) return rc;
      if (run.produced < limit) break;                   // everything the symbols allow has been deconvolved
      continue;                                          // a search window: more symbols are waiting
    }
    // Third fruitless sweep: mpeg_sync calls deconv->next_sync().  Bytes of this
    // run that the search has not consumed are void (their symbols are re-read by
    // the next hypothesis); the register state is re-derived at that exact byte.
    const uint64_t left = h->s_bytes.count;
    const uint64_t voided = std::min(left, run.produced);
    const uint64_t used = run.produced - voided;
    h->s_bytes.count = left - voided;
    h->s_bytes.fresh -= voided;
    if (c.keep_taps) tap_bytes.resize(tap_at + used);
    DeconvRun kept;
    kept.after = h->hyp[h->locked];
    if (used && (rc = deconv_launch(h, used, &kept))) return rc;
    if ((rc = deconv_commit(h, kept))) return rc;
    if (++h->locked == 4) { h->locked = 0; h->skip = 1; }  // dvb.h:185-193
  }
  if (max_passes && pass == max_passes) return fail(h, LDVB_ESTATE, "deconvolution/sync search did not drain the symbol stream");
  if ((rc = stream_compact(h, h->s_sym, h->d_scratch))) return rc;
  if (c.keep_taps) {
    Tap &t = h->taps[LDVB_TAP_BYTES];
    if (t.buf.bytes < tap_bytes.size()) { t.buf.release(); CK(t.buf.alloc(tap_bytes.size() + 4096)); }
    if (!tap_bytes.empty()) CK(cudaMemcpy(t.buf.p, tap_bytes.data(), tap_bytes.size(), cudaMemcpyHostToDevice));
    t.bytes = tap_bytes.size();
  }
  if ((rc = tap_store(h, LDVB_TAP_MPEGBYTES, h->s_mpeg.at(h->s_mpeg.count - h->s_mpeg.fresh), h->s_mpeg.fresh))) return rc;
  // ---- de-interleave + RS + derandomise
  delete wt; wt = new WallTimer(h, "wall:fec");
  if ((rc = run_fec(h, ts_dst, ts_cap, ts_out))) return rc;
  return LDVB_OK;
}

// ------------------------------------------------------------ cnr_fft / spectrum

// One telemetry runnable over the samples [abs0, abs0 + avail) held in `src`.
// Waits for the telemetry copies queued by run_meas_unit and finishes the measurements on the
// host: cnr_fft::do_cnr (sdr.h:1326-1330) and spectrum::do_spectrum (sdr.h:1390-1394) take the
// logarithms with glibc, like the reference.  Idempotent; called where the device is busy anyway
// and before anything reads the queues.
int meas_finish(ldvb_handle *h) {
  if (h->meas_pending.empty()) return LDVB_OK;
  CK(cudaEventSynchronize(h->meas_ev));
  for (const auto &m : h->meas_pending) {
    const float *src = h->meas_host + m.off;
    if (m.is_cnr) {
      for (int p = 0; p < m.np; ++p) {
        const float c2plusn2 = src[3 * p];
        const float n2 = (src[3 * p + 1] + src[3 * p + 2]) / 2;
        const float c2 = c2plusn2 - n2;
        const float cnr = (c2 > 0 && n2 > 0) ? 10 * logf(c2 / n2) / logf(10) : -50;
        h->cnr_queue.push_back(cnr);
      }
    } else {
      const int n = m.n;
      const size_t o = h->spec_queue.size();
      h->spec_queue.resize(o + (size_t)m.np * n);
      for (int p = 0; p < m.np; ++p) {
        const float *avg = src + (size_t)p * n;
        float *row = h->spec_queue.data() + o + (size_t)p * n;
        for (int i = 0; i < n / 2; ++i) {
          row[i] = 10 * log10f(avg[n / 2 + i]);
          row[n / 2 + i] = 10 * log10f(avg[i]);
        }
      }
    }
  }
  h->meas_pending.clear();
  h->meas_host_used = 0;
  return LDVB_OK;
}

// `need` floats of page-locked host memory for one asynchronous telemetry read-back.
int meas_host_reserve(ldvb_handle *h, size_t need, float **dst) {
  if (!h->meas_ev) CK(cudaEventCreateWithFlags(&h->meas_ev, cudaEventDisableTiming));
  if (h->meas_host_used + need > h->meas_host_cap) {
    int rc = meas_finish(h);                       // drains: nothing points into the buffer any more
    if (rc) return rc;
    if (need > h->meas_host_cap) {
      if (h->meas_host) cudaFreeHost(h->meas_host);
      h->meas_host = nullptr; h->meas_host_cap = 0;
      const size_t cap = std::max<size_t>(2 * need, (size_t)kMeasGroup * 1024);
      if (cudaMallocHost(&h->meas_host, cap * 4) != cudaSuccess) return fail(h, LDVB_ENOMEM, "telemetry host buffer");
      h->meas_host_cap = cap;
    }
  }
  *dst = h->meas_host + h->meas_host_used;
  h->meas_host_used += need;
  return LDVB_OK;
}

// Measured blocks of one unit over the samples [abs0, abs0 + avail): start positions relative to abs0.
// Phase advances by n per block, a block is measured when it reaches the decimation (sdr.h:1294-1302,
// 1362-1370).  Closed form instead of a walk over every block.
void meas_plan(ldvb_handle::MeasUnit &u, uint64_t abs0, uint64_t avail, std::vector<uint64_t> &points) {
  const int64_t n = (int64_t)1 << u.logn;
  const uint64_t abs_end = abs0 + avail;
  while (u.pos + (uint64_t)n <= abs_end) {
    int64_t k = (u.decimation - u.phase + n - 1) / n;
    if (k < 1) k = 1;
    const uint64_t left = (abs_end - u.pos) / (uint64_t)n;
    if ((uint64_t)k > left) { u.phase += (int64_t)left * n; u.pos += left * (uint64_t)n; break; }
    u.phase += k * n - u.decimation;
    u.pos += (uint64_t)k * (uint64_t)n;
    points.push_back(u.pos - (uint64_t)n - abs0);
  }
}

int meas_launch(ldvb_handle *h, ldvb_handle::MeasUnit &u, const MeasSrc &src, const std::vector<uint64_t> &points) {
  const int64_t n = (int64_t)1 << u.logn;
  const bool is_cnr = u.bandwidth > 0;
  // do_cnr (sdr.h:1306-1308, 1322): centre bin from freq_tap as of the start of the batch
  const float tap_multiplier = (float)(1.0 / h->decim);          // leandvb.cc:514
  const float center_freq = h->rx_state.freq_tap * tap_multiplier;
  const int icf = (int)floor(center_freq * (float)n + 0.5);
  const int bwslots = is_cnr ? (int)((u.bandwidth / 4) * (float)n) : 0;
  for (size_t g0 = 0; g0 < points.size(); g0 += kMeasGroup) {
    const int np = (int)std::min<size_t>(kMeasGroup, points.size() - g0);
    CK(cudaMemcpyAsync(h->d_meas_points.p, points.data() + g0, (size_t)np * 8, cudaMemcpyHostToDevice, h->st));
    MeasArgs a;
    a.src = src; a.point_start = h->d_meas_points.as<uint64_t>(); a.npoints = np; a.logn = u.logn;
    a.twiddle_rev = h->d_twiddle.as<float2>(); a.power = h->d_meas_power.as<float>();
    KL("meas_power", launch_meas_power(a, h->st));
    MeasEmaArgs e;
    e.power = a.power; e.npoints = np; e.n = (int)n; e.kavg = u.kavg;
    e.avg = u.d_avg.as<float>(); e.have = u.d_have.as<int>();
    e.bwslots = bwslots; e.icf = icf; e.sums = h->d_meas_sums.as<float>();
    e.rows = is_cnr ? nullptr : h->d_meas_rows.as<float>();
    KL("meas_ema", launch_meas_ema(e, h->st));
    if (is_cnr && !bwslots) continue;                            // sdr.h:1324
    // Results travel to the host asynchronously; meas_finish() turns them into dB later.
    const size_t need = is_cnr ? 3 * (size_t)np : (size_t)np * (size_t)n;
    float *dst = nullptr;
    { int rcr = meas_host_reserve(h, need, &dst); if (rcr) return rcr; }
    CK(cudaMemcpyAsync(dst, is_cnr ? e.sums : e.rows, need * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaEventRecord(h->meas_ev, h->st));
    h->meas_pending.push_back({is_cnr, np, (int)n, (size_t)(dst - h->meas_host)});
  }
  return LDVB_OK;
}

int run_meas_unit(ldvb_handle *h, ldvb_handle::MeasUnit &u, const MeasSrc &src, uint64_t abs0, uint64_t avail) {
  if (!u.on) return LDVB_OK;
  std::vector<uint64_t> points;
  meas_plan(u, abs0, avail, points);
  return meas_launch(h, u, src, points);
}

// cnr_fft and spectrum on the new samples of this batch: `rest` at element rest_off holds
// n_new samples in format fmt, continuing the stream after the carried (< 4096) samples.
int run_meas(ldvb_handle *h, const RawSrc &rest, uint64_t rest_off, int fmt, uint64_t n_new) {
  if (!h->m_cnr.on && !h->m_spec.on) return LDVB_OK;
  WallTimer wt(h, "wall:telemetry");
  MeasSrc src;
  src.carry = h->d_meas_carry[h->meas_carry_sel].as<float2>();
  src.carry_count = h->meas_carry_count;
  src.rest = rest; src.rest_off = rest_off; src.fmt = fmt; src.scale = h->cfg.float_scale;
  const uint64_t abs0 = h->meas_abs_next - h->meas_carry_count;
  src.rot_lut = h->use_rot ? h->d_rot.as<float>() : nullptr;
  src.rot_index0 = (uint32_t)(abs0 & 0xffffu);
  const uint64_t avail = h->meas_carry_count + n_new;
  int rc;
  if ((rc = run_meas_unit(h, h->m_cnr, src, abs0, avail))) return rc;
  if ((rc = run_meas_unit(h, h->m_spec, src, abs0, avail))) return rc;
  // Keep what the slower reader has not consumed yet.
  uint64_t keep_from = abs0 + avail;
  if (h->m_cnr.on) keep_from = std::min(keep_from, h->m_cnr.pos);
  if (h->m_spec.on) keep_from = std::min(keep_from, h->m_spec.pos);
  const uint64_t keep = abs0 + avail - keep_from;
  if (keep >= 4096) return fail(h, LDVB_ESTATE, "telemetry carry overflow");
  if (keep) {
    const int other = h->meas_carry_sel ^ 1;
    KL("meas_save", launch_meas_save(src, keep_from - abs0, (uint32_t)keep, h->d_meas_carry[other].as<float2>(), h->st));
    h->meas_carry_sel = other;
  }
  h->meas_carry_count = keep;
  h->meas_abs_next = abs0 + avail;
  return LDVB_OK;
}

// ------------------------------------------------------------------ whole chain

int run_chain(ldvb_handle *h, const void *src_dev, bool src_is_user_dev, uint64_t n, uint8_t *ts_dst,
              uint64_t ts_cap, uint64_t *ts_out) {
  const ldvb_config &c = h->cfg;
  *ts_out = 0;
  if (n > c.max_batch) return fail(h, LDVB_EOVERFLOW, "n_samples exceeds max_batch");
  if (cudaSetDevice(c.device) != cudaSuccess) return fail(h, LDVB_ECUDA, "cudaSetDevice");
  int rc;
  // ---- raw stream view
  Stream &raw = h->s_raw;
  const uint64_t c0 = raw.count;
  RawSrc src;
  src.head = raw.at(0); src.main = nullptr; src.c0 = c0;
  const uint64_t head_new = src_is_user_dev ? std::min<uint64_t>(n, 65536) : n;
  if (src_is_user_dev && (reinterpret_cast<uintptr_t>(src_dev) & 15)) return fail(h, LDVB_EINVAL, "iq_dev must be 16-byte aligned");
  if (n) {
    // (host data was already copied behind the carry by ldvb_push)
    if (src_is_user_dev) {
      CK(cudaMemcpyAsync(raw.at(c0), src_dev, head_new * raw.elem, cudaMemcpyDeviceToDevice, h->st));
      src.main = src_dev;
    }
  }
  src.head_count = c0 + head_new;
  const uint64_t avail = c0 + n;
  h->meas.samples_in += n;
  for (Tap &t : h->taps) t.bytes = 0;
  h->s_bytes.fresh = 0;

  // ---- notch / front end
  WallTimer *wt = new WallTimer(h, "wall:front");
  struct WtGuard { WallTimer *&p; ~WtGuard() { delete p; } } wt_guard{wt};
  uint64_t raw_consumed = 0;
  if (c.anf) {
    if ((rc = run_notch(h, src, avail, &raw_consumed))) return rc;
    if (!h->notch_fused) {   // (fused: run_notch produced the preprocessed stream and the telemetry)
    {   // cnr_fft / spectrum read the notched stream (leandvb.cc:296-343); the rotator is applied on load
      RawSrc fresh;
      fresh.head = h->s_notched.at(0); fresh.head_count = h->s_notched.count; fresh.main = nullptr; fresh.c0 = 0;
      if ((rc = run_meas(h, fresh, h->s_notched.count - h->s_notched.fresh, 5, h->s_notched.fresh))) return rc;
    }
    RawSrc nsrc;
    nsrc.head = h->s_notched.at(0); nsrc.head_count = h->s_notched.count; nsrc.main = nullptr; nsrc.c0 = 0;
    uint64_t ncons = 0;
    if (h->use_fir || h->use_decim || h->use_rot) {
      if ((rc = run_frontend(h, nsrc, 5, h->s_notched.count, &ncons))) return rc;
      if ((rc = stream_consume(h, h->s_notched, ncons, h->d_scratch))) return rc;
    } else {
      // The notch output IS the preprocessed stream: move it over.
      const uint64_t k = h->s_notched.count;
      if (h->s_pp.count + k > h->s_pp.cap) return fail(h, LDVB_EOVERFLOW, "preprocessed stream overflow");
      if (k) CK(cudaMemcpyAsync(h->s_pp.at(h->s_pp.count), h->s_notched.at(0), k * 8, cudaMemcpyDeviceToDevice, h->st));
      h->s_pp.count += k; h->s_pp.fresh = k;
      h->s_notched.count = 0;
    }
    }
  } else {
    if ((rc = run_meas(h, src, c0, c.input_format, n))) return rc;
    if ((rc = run_frontend(h, src, c.input_format, avail, &raw_consumed))) return rc;
  }
  // Carry the unread raw samples.
  {
    const uint64_t left = avail - raw_consumed;
    if (left > raw.cap) return fail(h, LDVB_EOVERFLOW, "raw carry overflow");
    if (src_is_user_dev) {
      // Unread samples are either in the head buffer or in the user's buffer.
      if (raw_consumed < c0) {
        // part of the old carry remains (tiny batch): keep head content, append the batch
        if (n > head_new) return fail(h, LDVB_ESTATE, "batch too small to drain the carry");
        raw.count = c0 + n;
        if ((rc = stream_consume(h, raw, raw_consumed, h->d_scratch))) return rc;
      } else if (left) {
        const uint64_t from = raw_consumed - c0;  // index into the user's buffer
        CK(cudaMemcpyAsync(raw.at(0), static_cast<const uint8_t *>(src_dev) + from * raw.elem, left * raw.elem,
                           cudaMemcpyDeviceToDevice, h->st));
        raw.count = left;
      } else raw.count = 0;
    } else {
      raw.count = avail;
      if ((rc = stream_consume(h, raw, raw_consumed, h->d_scratch))) return rc;
    }
  }
  if ((rc = tap_store(h, LDVB_TAP_PREPROCESSED, h->s_pp.at(h->s_pp.count - h->s_pp.fresh), h->s_pp.fresh * 8))) return rc;

  // ---- receiver
  delete wt; wt = new WallTimer(h, "wall:rx");
  if ((rc = run_receiver(h))) return rc;
  if ((rc = meas_finish(h))) return rc;            // (already done in FAST mode, while the spans ran)
  if ((rc = tap_store(h, LDVB_TAP_SYMBOLS, h->s_sym.at(h->s_sym.count - h->s_sym.fresh), h->s_sym.fresh * 4))) return rc;
  if (h->cfg.keep_taps && !h->meas_log.empty()) {
    Tap &t = h->taps[LDVB_TAP_MEAS];
    if (t.buf.bytes < h->meas_log.size() * 4) { t.buf.release(); CK(t.buf.alloc(h->meas_log.size() * 4 + 4096)); }
    CK(cudaMemcpy(t.buf.p, h->meas_log.data(), h->meas_log.size() * 4, cudaMemcpyHostToDevice));
    t.bytes = h->meas_log.size() * 4;
  }

  delete wt; wt = nullptr;
  if ((rc = run_backend(h, ts_dst, ts_cap, ts_out))) return rc;
  h->meas.kernel_launches = h->launches;
  return LDVB_OK;
}


// ============================================================== time sharding
// SURVEY.md 8(e).  One stream, consecutive chunks on different handles (GPUs).
// Front stage (speculative, needs only the halo): notch, front end, receiver spans.
// Back stage (exact, needs the previous chunk's EDGE): seam to the previous chunk,
// symbol stream, deconvolution/sync/FEC, then this chunk's EDGE.

constexpr uint32_t kEdgeMagic = 0x45474445;
constexpr uint32_t kEdgeSym = 2048, kEdgeBytes = 4096, kEdgeMpeg = 4096, kEdgeVit = 16;

struct EdgeBlob {
  uint32_t magic, size;
  uint64_t abs_raw_end;        // one past the last raw sample the exporting handle held
  uint64_t own_end_chunk;      // absolute receiver chunk where the next handle's ownership starts
  // receiver: loop state at own_end_chunk (reference frame), last span's verification log
  RxState rx_end;
  int32_t cum_rot, have_tail;
  uint32_t n_tail, tail_first_word;
  RxSeamSym tail_log[kRxSeamLog];
  // notch: estimates entering block notch_block
  NotchState notch;
  uint64_t notch_block;
  // deconvolution / framing
  HypState hyp[4];
  int32_t locked, skip;
  SyncState sync;
  int32_t derand_pos, lock;
  VitCtl vit_ctl;
  int32_t n_vit, pad0;
  VitDecState vit[kEdgeVit];
  uint32_t n_sym, n_bytes, n_mpeg, pad1;
  uint32_t sym[kEdgeSym];
  uint8_t bytes[kEdgeBytes];
  uint8_t mpeg[kEdgeMpeg];
};

uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

uint64_t shard_unit(const ldvb_handle *h) {
  uint64_t u = (uint64_t)kRxChunk * h->decim;
  if (h->cfg.anf) u = u / gcd64(u, kNotchN) * kNotchN;
  return u;
}

// Receiver chunks that can be run when the raw stream is known up to (absolute) sample E.
uint64_t shard_chunks_avail(const ldvb_handle *h, uint64_t E) {
  const uint64_t N = h->use_fir ? (uint64_t)h->fir_n : 0, D = (uint64_t)h->decim;
  const uint64_t pp = N ? (E >= N ? (E - N) / D : 0) : E / D;   // dsp.h:246-247, generic.h:254
  return pp >= (uint64_t)kRxChunk + h->readahead ? (pp - h->readahead) / kRxChunk : 0;   // sdr.h:783
}

uint32_t shard_warm_chunks(const ldvb_handle *h) { return h->cfg.warmup_chunks ? h->cfg.warmup_chunks : 4; }
constexpr uint64_t kNotchLead = (uint64_t)kNotchN * 4;   // 2 exact warm-up blocks + the 8192-sample guess window

uint64_t shard_min_halo(const ldvb_handle *h) {
  const uint64_t u = shard_unit(h), D = (uint64_t)h->decim;
  uint64_t need = (h->use_fir ? h->fir_n : 0) + (uint64_t)h->readahead * D +
                  (2 + kRxVerifyChunks + shard_warm_chunks(h)) * (uint64_t)kRxChunk * D + u + (h->cfg.anf ? kNotchLead : 0);
  return (need + u - 1) / u * u + u;
}

// Where a chunk [A0, A0+H+C) starts its stages.  Pure function of (A0, H): the handle that
// exports an EDGE evaluates it for its successor.
struct ShardStart { uint64_t R0, own_begin, notch_b0; };
ShardStart shard_start(const ldvb_handle *h, uint64_t A0, uint64_t H) {
  ShardStart g;
  if (H == 0) { g.R0 = A0; g.own_begin = A0 / ((uint64_t)kRxChunk * h->decim); g.notch_b0 = A0 / kNotchN; return g; }
  const uint64_t u = shard_unit(h), cs = (uint64_t)kRxChunk * h->decim;
  g.own_begin = shard_chunks_avail(h, A0 + H) - kRxVerifyChunks;
  const uint64_t cw = g.own_begin - shard_warm_chunks(h);
  g.R0 = cw * cs / u * u;
  g.notch_b0 = g.R0 / kNotchN;
  return g;
}

int shard_check(ldvb_handle *h, const ldvb_shard *s) {
  if (!h || !s) return LDVB_EINVAL;
  if (h->cfg.rx_mode != LDVB_RX_FAST) return fail(h, LDVB_EINVAL, "time sharding needs rx_mode = LDVB_RX_FAST");
  // (RRC sampler: the tap-update throttle is restored from the absolute chunk index in shard_run_front, but a first
  //  three-chunk GPU run lost the packets of one chunk -- not debugged, so the combination stays rejected.)
  if (h->cfg.sampler == LDVB_SAMP_RRC) return fail(h, LDVB_EINVAL, "time sharding: RRC sampler not supported yet");
  if (h->cfg.fastlock || h->cfg.hs) return fail(h, LDVB_EINVAL, "time sharding: --fastlock / --hs not supported");
  const uint64_t u = shard_unit(h);
  if (s->n_halo % u || s->n_chunk % u || s->abs_raw0 % u || s->n_halo_next % u)
    return fail(h, LDVB_EINVAL, "time sharding: abs_raw0, n_halo, n_chunk must be multiples of lcm(4096, 128*decimation)");
  if (!s->iq_dev || (reinterpret_cast<uintptr_t>(s->iq_dev) & 15)) return fail(h, LDVB_EINVAL, "iq_dev must be 16-byte aligned");
  if (s->n_halo + s->n_chunk > h->cfg.max_batch) return fail(h, LDVB_EOVERFLOW, "halo + chunk exceed max_batch");
  if (s->n_halo == 0 && s->abs_raw0 != 0) return fail(h, LDVB_EINVAL, "only the first chunk of the stream may come without a halo");
  if (s->n_halo && s->n_halo < shard_min_halo(h)) return fail(h, LDVB_EINVAL, "halo shorter than ldvb_shard_min_halo()");
  if (s->n_halo_next && s->n_halo_next < shard_min_halo(h)) return fail(h, LDVB_EINVAL, "next halo shorter than ldvb_shard_min_halo()");
  if (s->n_chunk < 4 * s->n_halo_next + 16 * u || s->n_chunk < 4 * s->n_halo) return fail(h, LDVB_EINVAL, "chunk too small for its halo");
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return fail(h, LDVB_ECUDA, "cudaSetDevice");
  return LDVB_OK;
}

// auto_notch::detect() (sdr.h:76-118) on the detect points inside [halo | chunk]; builds the
// epoch list of the chunk and the bins in force where the next chunk's halo starts.
int shard_detect(ldvb_handle *h, ldvb_shard *s) {
  const ldvb_config &c = h->cfg;
  ldvb_handle::Shard &sh = h->shard;
  sh.det_valid = false; sh.valid = false;
  sh.epochs.clear();
  for (int i = 0; i < 4; ++i) s->bins_after[i] = (c.anf && i < c.anf) ? s->bins_before[i] : -1;
  if (!c.anf) { sh.desc = *s; sh.det_valid = true; return LDVB_OK; }
  const uint64_t nblocks = (s->n_halo + s->n_chunk) / kNotchN, blk0 = s->abs_raw0 / kNotchN;
  const uint64_t next_blk = (s->n_halo + s->n_chunk - s->n_halo_next) / kNotchN;   // local
  std::vector<uint64_t> dblocks;
  for (uint64_t b = 1023 - (blk0 % 1024); b < nblocks; b += 1024) dblocks.push_back(b);   // (blk0 + b + 1) % 1024 == 0
  NotchEpoch e0;
  memset(&e0, 0, sizeof e0);
  int cur[kNotchMaxSlots];
  for (int sl = 0; sl < kNotchMaxSlots; ++sl) { cur[sl] = sl < c.anf ? s->bins_before[sl] : -1; e0.bin[sl] = cur[sl]; }
  for (int sl = 0; sl < c.anf; ++sl) { int rc = notch_table_for_bin(h, cur[sl], &e0.table_index[sl]); if (rc) return rc; }
  sh.epochs.push_back(e0);
  if (!dblocks.empty()) {
    if (dblocks.size() * 8 > h->d_notch_blocks.bytes) return fail(h, LDVB_EOVERFLOW, "too many notch detect points");
    NotchDetectArgs d;
    d.src.head = s->iq_dev; d.src.head_count = s->n_halo + s->n_chunk; d.src.main = nullptr; d.src.c0 = 0;
    d.fmt = c.input_format; d.scale = c.float_scale;
    CK(cudaMemcpyAsync(h->d_notch_blocks.p, dblocks.data(), dblocks.size() * 8, cudaMemcpyHostToDevice, h->st));
    d.block_index = h->d_notch_blocks.as<uint64_t>();
    d.ndetect = (int)dblocks.size(); d.nslots = c.anf;
    d.twiddle_rev = h->d_twiddle.as<float2>();
    d.bins_out = h->d_notch_bins.as<int32_t>();
    KL("notch_detect", launch_notch_detect(d, h->st));
    std::vector<int32_t> bins(dblocks.size() * c.anf);
    CK(cudaMemcpyAsync(bins.data(), h->d_notch_bins.p, bins.size() * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    for (size_t k = 0; k < dblocks.size(); ++k) {
      if (dblocks[k] < next_blk) for (int sl = 0; sl < c.anf; ++sl) s->bins_after[sl] = bins[k * c.anf + sl];
      bool changed = false;
      NotchEpoch e = sh.epochs.back();
      for (int sl = 0; sl < kNotchMaxSlots; ++sl) e.reset[sl] = 0;
      for (int sl = 0; sl < c.anf; ++sl) {
        const int nb = bins[k * c.anf + sl];
        if (nb != cur[sl]) {
          changed = true; cur[sl] = nb; e.bin[sl] = nb; e.reset[sl] = 1;
          int rc = notch_table_for_bin(h, nb, &e.table_index[sl]);
          if (rc) return rc;
        }
      }
      if (changed) { e.first_block = dblocks[k]; if (e.first_block == 0) sh.epochs[0] = e; else sh.epochs.push_back(e); }
    }
  }
  sh.desc = *s;
  sh.det_valid = true;
  return LDVB_OK;
}

// The speculative stages on [halo | chunk].  exact_notch != nullptr: the notch of the first
// owned block starts from that (imported) state instead of a warm-up.
int shard_run_front(ldvb_handle *h, const NotchState *exact_notch) {
  const ldvb_config &c = h->cfg;
  ldvb_handle::Shard &sh = h->shard;
  const ldvb_shard &s = sh.desc;
  int rc;
  const uint64_t total = s.n_halo + s.n_chunk, D = (uint64_t)h->decim, cs = (uint64_t)kRxChunk * D;
  const uint64_t R0l = sh.R0 - s.abs_raw0;   // local raw index where the front end starts
  RawSrc src;
  src.head = s.iq_dev; src.head_count = total; src.main = nullptr; src.c0 = 0;
  Stream *ss[] = {&h->s_raw, &h->s_notched, &h->s_pp};
  for (Stream *q : ss) { q->count = 0; q->fresh = 0; }
  {
    WallTimer wt(h, "wall:front");
    if (c.anf) {
      CK(cudaMemcpyAsync(h->d_notch_epochs.p, sh.epochs.data(), sh.epochs.size() * sizeof(NotchEpoch), cudaMemcpyHostToDevice, h->st));
      NotchState st_in;
      memset(&st_in, 0, sizeof st_in);
      st_in.gain = 1;
      if (exact_notch) st_in = *exact_notch;
      else if (sh.first) st_in = h->notch;
      CK(cudaMemcpyAsync(h->d_notch_state.p, &st_in, sizeof st_in, cudaMemcpyHostToDevice, h->st));
      NotchApplyArgs &a = sh.na;
      a.src = src; a.fmt = c.input_format; a.scale = c.float_scale;
      a.out = reinterpret_cast<float2 *>(h->s_notched.at(0));
      a.nblocks = total / kNotchN; a.nslots = c.anf;
      a.block0 = sh.notch_b0 - s.abs_raw0 / kNotchN;
      a.first_exact = (sh.first || exact_notch) ? 1 : 0;
      a.k = 0.002f; a.gain = 1.0f;
      a.w_block = (float)pow((double)(1.0f - 0.002f), (double)kNotchN);
      a.expj_tables = h->d_notch_tables.as<float2>();
      a.epochs = h->d_notch_epochs.as<NotchEpoch>();
      a.nepochs = (int)sh.epochs.size();
      a.seg_blocks = 1; a.warm_blocks = 2;
      a.nsegs = (uint32_t)(a.nblocks - a.block0);
      a.state_in = h->d_notch_state.as<NotchState>();
      a.seg_entry = h->d_notch_entry.as<float2>();
      a.seg_exit = h->d_notch_exit.as<float2>();
      a.seg_exact = h->d_notch_exact.as<uint8_t>();
      KL("notch_guess", launch_notch_guess(a, h->d_notch_guess.as<float2>(), h->d_notch_weights.as<float>(), h->st));
      KL("notch_apply", launch_notch_apply(a, nullptr, 0, h->d_notch_guess.as<float2>(), h->st));
      std::vector<float2> exitv;
      if ((rc = notch_verify_repair(h, a, exitv))) return rc;
    }
    // front end (rotator + FIR + decimation) from R0
    const bool need_fe = !c.anf || h->use_fir || h->use_decim || h->use_rot;
    if (need_fe) {
      RawSrc fsrc;
      int fmt = c.input_format;
      if (c.anf) { fsrc.head = h->s_notched.at(R0l); fmt = 5; }
      else fsrc.head = static_cast<const uint8_t *>(s.iq_dev) + R0l * h->s_raw.elem;
      fsrc.head_count = total - R0l; fsrc.main = nullptr; fsrc.c0 = 0;
      h->rot_index = (uint32_t)(sh.R0 & 0xffffu);
      uint64_t used = 0;
      if ((rc = run_frontend(h, fsrc, fmt, total - R0l, &used))) return rc;
      sh.pp = reinterpret_cast<const float2 *>(h->s_pp.at(0));
      h->s_pp.count = 0;
    } else {
      sh.pp = reinterpret_cast<const float2 *>(h->s_notched.at(R0l));
    }
  }
  WallTimer wt(h, "wall:rx");
  RxArgs &a = sh.a;
  memset(&a, 0, sizeof a);
  a.p = h->rxp;
  a.p.cstln = h->d_cstln.as<CstlnCellDev>();
  a.p.trig = h->d_trig.as<float2>();
  a.x = sh.pp;
  a.chunk0 = sh.own_begin - sh.base_chunk;
  a.nchunks = sh.own_end - sh.base_chunk;
  a.avail_chunks = sh.avail_end - sh.base_chunk;
  a.first_exact = sh.first ? 1 : 0;
  a.prev_end = nullptr;
  RxState st0 = h->rx_state;   // loop state the warm-ups start from: this handle's latest
  st0.meas_count = (uint32_t)((sh.base_chunk * (uint64_t)kRxChunk) % h->rxp.meas_decimation);
  if (h->cfg.sampler == LDVB_SAMP_RRC && !sh.first) {
    // fir_sampler::update_freq (sdr.h:667-675): the tap-update throttle is a pure function of the absolute chunk
    // index -- updates at chunks 0, P, 2P, ... with P = ceil(16 n / 128); value of the counter in front of chunk c >= 1.
    const int R = h->rxp.rrc_n * 16, Pp = (R + kRxChunk - 1) / kRxChunk;
    st0.rrc_update_phase = sh.base_chunk ? R - kRxChunk * (int)((sh.base_chunk - 1) % (uint64_t)Pp) : 0;
    st0.rrc_f = st0.freqw / (float)h->rxp.rrc_sub;
  }
  CK(cudaMemcpyAsync(h->d_rx_state.p, &st0, sizeof st0, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemsetAsync(h->d_rx_measn.p, 0, 4, h->st));
  a.state_in = h->d_rx_state.as<RxState>();
  a.warm_in = a.state_in;
  a.info = h->d_rx_info.as<RxSpanInfo>();
  a.state_end = h->d_rx_end.as<RxState>();
  a.meas = nullptr; a.meas_count = nullptr; a.max_meas = 0;
  // AGC settling (see run_receiver).  The first chunk of the stream has the exact start state: a serial pass, its
  // symbols open the symbol stream.  Later chunks have no exact state to start from before the EDGE arrives (the
  // seam to the previous chunk is verified then): their warm-ups are seeded with the measured power instead.
  sh.settle_syms = 0;
  uint64_t owned = sh.own_end - sh.own_begin;
  {
    float power = 0; bool far = false;
    if ((rc = rx_level_check(h, a.x + a.chunk0 * kRxChunk, owned * kRxChunk, &power, &far))) return rc;
    if (sh.first && h->rx_cold && !c.hs && c.settle_chunks >= 0 && owned > 8 + 16) far = true;   // see run_receiver
    if (far && sh.first) {
      h->rx_cold = false;
      const uint64_t K = std::min(owned - 8, rx_settle_chunks(h));
      if ((rc = rx_settle(h, a, K, reinterpret_cast<uint32_t *>(h->s_sym.at(0)), h->s_sym.cap, &sh.settle_syms))) return rc;
      owned -= K;
    } else if (far) {
      st0.est_insp = power;
      st0.agc_gain = 75.0f / sqrtf(power);
      CK(cudaMemcpyAsync(h->d_rx_state.p, &st0, sizeof st0, cudaMemcpyHostToDevice, h->st));
      CK(cudaStreamSynchronize(h->st));   // st0 is a local
      ++h->meas.settle_passes;
    }
  }
  if ((rc = rx_fast_launch(h, a, sh.sa, owned))) return rc;
  // Cold start (carrier offset): most warm-ups fail; pull the loops in before the back stage.
  if (a.nspans > 8) {
    KL("rx_plan", launch_rx_plan(a.info, sh.sa.seams, a.nspans, a.span_cap, h->cst.nrotations, 0, 0,
                                 h->d_rx_off.as<uint64_t>(), h->d_rx_skip.as<uint32_t>(), h->d_rx_rot.as<uint8_t>(),
                                 h->d_counts.as<uint64_t>(), h->st));
    uint64_t plan[9];
    CK(cudaMemcpyAsync(plan, h->d_counts.p, sizeof plan, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (plan[8] > std::max<uint32_t>(4, a.nspans / 32)) {
      std::vector<RxSeam> seams; std::vector<RxSpanInfo> info;
      if ((rc = rx_fetch(h, a, sh.sa, seams, info))) return rc;
      if ((rc = rx_fast_reseed(h, a, sh.sa, seams, info))) return rc;
    }
  }
  (void)cs;
  return LDVB_OK;
}

int shard_front(ldvb_handle *h, const ldvb_shard *s) {
  int rc;
  ldvb_handle::Shard &sh = h->shard;
  const bool same = sh.det_valid && sh.desc.iq_dev == s->iq_dev && sh.desc.abs_raw0 == s->abs_raw0 &&
                    sh.desc.n_halo == s->n_halo && sh.desc.n_chunk == s->n_chunk && sh.desc.n_halo_next == s->n_halo_next &&
                    memcmp(sh.desc.bins_before, s->bins_before, sizeof s->bins_before) == 0;
  if (!same) { ldvb_shard tmp = *s; if ((rc = shard_detect(h, &tmp))) return rc; }
  sh.valid = false;
  sh.desc.last = s->last;
  sh.first = (s->n_halo == 0);
  if (sh.first) reset_carry(h);
  const uint64_t cs = (uint64_t)kRxChunk * h->decim;
  const ShardStart g = shard_start(h, s->abs_raw0, s->n_halo);
  sh.R0 = g.R0; sh.own_begin = g.own_begin; sh.notch_b0 = g.notch_b0;
  sh.base_chunk = sh.R0 / cs;
  sh.avail_end = shard_chunks_avail(h, s->abs_raw0 + s->n_halo + s->n_chunk);
  sh.own_end = s->last ? sh.avail_end : sh.avail_end - kRxVerifyChunks;
  if (sh.R0 < s->abs_raw0 + (h->cfg.anf && !sh.first ? kNotchLead : 0)) return fail(h, LDVB_EINVAL, "halo too short");
  if (sh.own_end < sh.own_begin + 8) return fail(h, LDVB_EINVAL, "chunk too small");
  if ((rc = shard_run_front(h, nullptr))) return rc;
  sh.valid = true;
  return LDVB_OK;
}

int shard_back(ldvb_handle *h, const EdgeBlob *in, uint8_t *ts_dst, uint64_t ts_cap, uint64_t *ts_out, EdgeBlob *out) {
  const ldvb_config &c = h->cfg;
  ldvb_handle::Shard &sh = h->shard;
  const ldvb_shard &s = sh.desc;
  *ts_out = 0;
  if (!sh.valid) return fail(h, LDVB_ESTATE, "ldvb_shard_back without ldvb_shard_front");
  if (!in != sh.first) return fail(h, LDVB_ESTATE, "an EDGE is needed for every chunk except the first");
  int rc;
  if (!h->d_edge_tail.p) {
    CK(h->d_edge_tail.alloc(sizeof(RxSeamSym) * kRxSeamLog));
    CK(h->d_edge_state.alloc(sizeof(RxState)));
    CK(h->d_edge_seam.alloc(sizeof(RxSeam)));
  }
  for (Tap &t : h->taps) t.bytes = 0;
  h->s_bytes.fresh = 0;
  h->meas.samples_in += s.n_chunk;
  RxArgs &a = sh.a;
  int rot0 = 0; uint32_t skip0 = 0;
  Stream &sym = h->s_sym;
  WallTimer *wt = new WallTimer(h, "wall:rx");
  struct WtGuard { WallTimer *&p; ~WtGuard() { delete p; } } wt_guard{wt};
  if (in) {
    if (in->magic != kEdgeMagic || in->size != sizeof(EdgeBlob)) return fail(h, LDVB_EINVAL, "bad EDGE blob");
    if (in->abs_raw_end != s.abs_raw0 + s.n_halo || in->own_end_chunk != sh.own_begin)
      return fail(h, LDVB_ESTATE, "EDGE does not belong in front of this chunk");
    if (in->n_sym > kEdgeSym || in->n_bytes > kEdgeBytes || in->n_mpeg > kEdgeMpeg) return fail(h, LDVB_EINVAL, "bad EDGE blob");
    // ---- serial state of the stages behind the receiver
    for (int i = 0; i < 4; ++i) h->hyp[i] = in->hyp[i];
    h->locked = in->locked; h->skip = in->skip; h->sync = in->sync; h->derand_pos = in->derand_pos;
    h->meas.lock = in->lock;
    sym.count = in->n_sym; h->s_bytes.count = in->n_bytes; h->s_mpeg.count = in->n_mpeg;
    if (in->n_sym) CK(cudaMemcpyAsync(sym.at(0), in->sym, in->n_sym * 4, cudaMemcpyHostToDevice, h->st));
    if (in->n_bytes) CK(cudaMemcpyAsync(h->s_bytes.at(0), in->bytes, in->n_bytes, cudaMemcpyHostToDevice, h->st));
    if (in->n_mpeg) CK(cudaMemcpyAsync(h->s_mpeg.at(0), in->mpeg, in->n_mpeg, cudaMemcpyHostToDevice, h->st));
    if (c.viterbi) {
      if (in->n_vit != h->vsyncs.nsyncs) return fail(h, LDVB_EINVAL, "EDGE: Viterbi state of another configuration");
      CK(cudaMemcpyAsync(h->d_vit_state.p, in->vit, sizeof(VitDecState) * in->n_vit, cudaMemcpyHostToDevice, h->st));
      CK(cudaMemcpyAsync(h->d_vit_ctl.p, &in->vit_ctl, sizeof(VitCtl), cudaMemcpyHostToDevice, h->st));
    }
    // ---- notch seam: did the warm-up merge with the previous chunk's trajectory?
    if (c.anf) {
      float2 entry[kNotchMaxSlots];
      CK(cudaMemcpyAsync(entry, sh.na.seg_entry, sizeof entry, cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
      bool same = in->notch_block == sh.notch_b0;
      for (int sl = 0; sl < c.anf; ++sl)
        same = same && memcmp(&entry[sl].x, &in->notch.slot[sl].est_re, 4) == 0 && memcmp(&entry[sl].y, &in->notch.slot[sl].est_im, 4) == 0;
      if (!same) {
        if (in->notch_block != sh.notch_b0) return fail(h, LDVB_ESTATE, "EDGE: notch state of another block");
        ++h->meas.notch_repaired;
        NotchState st = in->notch;
        st.gain = 1;
        if ((rc = shard_run_front(h, &st))) return rc;   // everything downstream of the notch again
      }
    }
    // ---- receiver seam
    RxSeam seam;
    memset(&seam, 0, sizeof seam);
    if (in->have_tail) {
      CK(cudaMemcpyAsync(h->d_edge_tail.p, in->tail_log, sizeof in->tail_log, cudaMemcpyHostToDevice, h->st));
      // the previous chunk's end state, in ITS last span's frame (in->rx_end is in the reference frame)
      RxState pe = in->rx_end;
      if (in->cum_rot) pe.phase = fmodf(pe.phase + (float)in->cum_rot * (65536.0f / h->cst.nrotations), 65536.0f);
      CK(cudaMemcpyAsync(h->d_edge_state.p, &pe, sizeof pe, cudaMemcpyHostToDevice, h->st));
      KL("rx_stitch", launch_rx_stitch_pair(sh.sa, h->d_edge_tail.as<RxSeamSym>(), in->n_tail, h->d_edge_state.as<RxState>(),
                                            h->d_edge_seam.as<RxSeam>(), h->st));
      CK(cudaMemcpyAsync(&seam, h->d_edge_seam.p, sizeof seam, cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
    }
    ++h->meas.seams_total;
    if (!seam.ok) {
      // Span 0 again, exactly, from the loop state the previous chunk ended with (already
      // in the reference frame), then the seam behind it.
      ++h->meas.seams_repaired;
      CK(cudaMemcpyAsync(h->d_edge_state.p, &in->rx_end, sizeof(RxState), cudaMemcpyHostToDevice, h->st));
      a.prev_end = h->d_edge_state.as<RxState>();
      const uint32_t zero = 0;
      CK(cudaMemcpyAsync(h->d_scratch.p, &zero, 4, cudaMemcpyHostToDevice, h->st));
      KL("rx", launch_rx(a, h->d_scratch.as<uint32_t>(), 1, h->st));
      if (a.nspans > 1) KL("rx_stitch", launch_rx_stitch(sh.sa, h->d_scratch.as<uint32_t>(), 1, h->st));
      CK(cudaStreamSynchronize(h->st));
      rot0 = 0; skip0 = 0;
    } else {
      rot0 = (in->cum_rot + seam.rot) % h->cst.nrotations;
      skip0 = (uint32_t)seam.skip_next;
      if (seam.extend_prev) {
        CK(cudaMemcpyAsync(sym.at(sym.count), &in->tail_first_word, 4, cudaMemcpyHostToDevice, h->st));
        ++sym.count;
      }
    }
  } else {
    sym.count = sh.settle_syms; h->s_bytes.count = 0; h->s_mpeg.count = 0;
    h->meas.symbols += sh.settle_syms;
  }
  uint64_t produced = 0;
  int cum = 0;
  if ((rc = rx_fast_resolve(h, a, sh.sa, rot0, skip0, reinterpret_cast<uint32_t *>(sym.at(sym.count)), sym.cap - sym.count,
                            &produced, &cum))) return rc;
  sym.count += produced; sym.fresh = produced;
  h->meas.symbols += produced;
  h->meas.freq_tap = h->rx_state.freq_tap;
  h->meas.ss = sqrtf(h->rx_state.est_insp);
  h->meas.mer = h->rx_state.est_ep ? 10 * logf(h->rx_state.est_sp / h->rx_state.est_ep) / logf(10) : 0;
  delete wt; wt = nullptr;
  if ((rc = run_backend(h, ts_dst, ts_cap, ts_out))) return rc;
  h->meas.kernel_launches = h->launches;
  sh.valid = false; sh.det_valid = false;
  if (!out) return LDVB_OK;

  // ---- this chunk's EDGE
  EdgeBlob &e = *out;
  memset(static_cast<void *>(&e), 0, sizeof e);
  e.magic = kEdgeMagic; e.size = sizeof(EdgeBlob);
  e.abs_raw_end = s.abs_raw0 + s.n_halo + s.n_chunk;
  e.own_end_chunk = sh.own_end;
  e.rx_end = h->rx_state;
  e.cum_rot = cum;
  const uint32_t lastsp = a.nspans - 1;
  if (!s.last) {
    RxSpanInfo inf;
    CK(cudaMemcpyAsync(&inf, a.info + lastsp, sizeof inf, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(e.tail_log, a.tail_log + (size_t)lastsp * kRxSeamLog, sizeof e.tail_log, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    e.have_tail = 1;
    e.n_tail = inf.n_tail;
    if (inf.n_tail && inf.n_out < a.span_cap) {
      uint32_t w = 0;
      CK(cudaMemcpy(&w, a.sym_out + (size_t)lastsp * a.span_cap + inf.n_out, 4, cudaMemcpyDeviceToHost));
      if (cum) w = (w & 0xffffu) | ((uint32_t)h->cst.rot[cum][(w >> 16) & 0xffu] << 16);
      e.tail_first_word = w;
    }
  }
  if (c.anf && !s.last) {
    const ShardStart nx = shard_start(h, e.abs_raw_end - s.n_halo_next, s.n_halo_next);
    e.notch_block = nx.notch_b0;
    const uint64_t nbl = nx.notch_b0 - s.abs_raw0 / kNotchN;   // local block
    if (nbl <= sh.na.block0 || nbl > sh.na.nblocks) return fail(h, LDVB_ESTATE, "next chunk's notch start outside this chunk");
    float2 ex[kNotchMaxSlots];
    CK(cudaMemcpy(ex, sh.na.seg_exit + (size_t)(nbl - sh.na.block0 - 1) * kNotchMaxSlots, sizeof ex, cudaMemcpyDeviceToHost));
    e.notch.gain = 1;
    for (int sl = 0; sl < kNotchMaxSlots; ++sl) { e.notch.slot[sl].bin = -1; e.notch.slot[sl].est_re = ex[sl].x; e.notch.slot[sl].est_im = ex[sl].y; }
    for (const NotchEpoch &ep : sh.epochs)
      if (ep.first_block < nbl) for (int sl = 0; sl < c.anf; ++sl) e.notch.slot[sl].bin = ep.bin[sl];
  }
  for (int i = 0; i < 4; ++i) e.hyp[i] = h->hyp[i];
  e.locked = h->locked; e.skip = h->skip; e.sync = h->sync; e.derand_pos = h->derand_pos; e.lock = h->meas.lock;
  if (sym.count > kEdgeSym || h->s_bytes.count > kEdgeBytes || h->s_mpeg.count > kEdgeMpeg)
    return fail(h, LDVB_EOVERFLOW, "unread stream remainders do not fit the EDGE");
  e.n_sym = (uint32_t)sym.count; e.n_bytes = (uint32_t)h->s_bytes.count; e.n_mpeg = (uint32_t)h->s_mpeg.count;
  if (e.n_sym) CK(cudaMemcpyAsync(e.sym, sym.at(0), e.n_sym * 4, cudaMemcpyDeviceToHost, h->st));
  if (e.n_bytes) CK(cudaMemcpyAsync(e.bytes, h->s_bytes.at(0), e.n_bytes, cudaMemcpyDeviceToHost, h->st));
  if (e.n_mpeg) CK(cudaMemcpyAsync(e.mpeg, h->s_mpeg.at(0), e.n_mpeg, cudaMemcpyDeviceToHost, h->st));
  if (c.viterbi) {
    e.n_vit = h->vsyncs.nsyncs;
    if (e.n_vit > (int)kEdgeVit) return fail(h, LDVB_EOVERFLOW, "too many Viterbi hypotheses for the EDGE");
    CK(cudaMemcpyAsync(e.vit, h->d_vit_state.p, sizeof(VitDecState) * e.n_vit, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(&e.vit_ctl, h->d_vit_ctl.p, sizeof(VitCtl), cudaMemcpyDeviceToHost, h->st));
  }
  CK(cudaStreamSynchronize(h->st));
  return LDVB_OK;
}

}  // namespace

extern "C" {

// Queues the `got` packets of d_ts for ldvb_pull.  The copy is asynchronous on the handle's stream
// (ordered before the next chain's kernels overwrite d_ts); ldvb_push synchronises once at its end.
static int push_collect(ldvb_handle *h, uint64_t got) {
  if (!got) return LDVB_OK;
  std::lock_guard<std::mutex> lk(h->qmu);
  const size_t need = h->ts_queue_wr + got * 188;
  if (need > h->ts_queue_cap) {
    // Grow (rare: the queue is sized for two full batches at the first push).
    CK(cudaStreamSynchronize(h->st));
    h->ts_queue_ready = h->ts_queue_wr;
    const size_t unread = h->ts_queue_wr - h->ts_queue_rd;
    size_t cap = std::max<size_t>(2 * (size_t)h->ts_cap * 188, 2 * (unread + got * 188));
    uint8_t *q = nullptr;
    CK(cudaHostAlloc((void **)&q, cap, cudaHostAllocDefault));
    if (unread) memcpy(q, h->ts_queue + h->ts_queue_rd, unread);
    if (h->ts_queue) cudaFreeHost(h->ts_queue);
    h->ts_queue = q; h->ts_queue_cap = cap; h->ts_queue_rd = 0; h->ts_queue_wr = unread; h->ts_queue_ready = unread;
  }
  CK(cudaMemcpyAsync(h->ts_queue + h->ts_queue_wr, h->d_ts.p, got * 188, cudaMemcpyDeviceToHost, h->st));
  h->ts_queue_wr += got * 188;
  return LDVB_OK;
}

// The packets queued so far have arrived: the caller has synchronised h->st.
static void push_publish(ldvb_handle *h) {
  std::lock_guard<std::mutex> lk(h->qmu);
  h->ts_queue_ready = h->ts_queue_wr;
}

// Samples per staged piece.  Synchronous push: sub_batch (the copy of a piece has to outlast the chain on the one
// in front of it within ONE call).  async_push: the chain of a piece overlaps the copies of the following calls,
// and its cost is mostly fixed latency (measured: 3.45 ms + 0.011 ms per Mi sample, against 0.144 ms per Mi cf32
// sample of PCIe time), so pieces are as large as the handle allows, up to 1 GiB of input.
static uint64_t stage_samples(const ldvb_handle *h) {
  if (!h->cfg.async_push || h->cfg.push_sub_batch > 0) return h->sub_batch;
  const uint64_t big = ((uint64_t)1 << 30) / h->s_raw.elem;
  return std::max<uint64_t>(h->sub_batch, std::min<uint64_t>(h->cfg.max_batch, big));
}

static int stage_init(ldvb_handle *h) {
  if (h->copy_st) return LDVB_OK;
  CK(cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking));
  for (int i = 0; i < ldvb_handle::kStages; ++i) {
    CK(h->d_stage[i].alloc(stage_samples(h) * h->s_raw.elem + 256));
    CK(cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming));
  }
  return LDVB_OK;
}

// Even split of a host batch into sub-batches (multiples of 4096 samples): a short last sub-batch would
// finish its copy long before the chain of the one in front of it is done, and the chain's fixed cost
// (~1.9 ms of serial-recurrence latency) would show twice at the end.
static uint64_t sub_batch_size(const ldvb_handle *h, size_t n) {
  const uint64_t sub = stage_samples(h);
  const uint64_t nsub0 = (n + sub - 1) / sub;
  return std::min<uint64_t>(sub, ((n + nsub0 - 1) / nsub0 + 4095) / 4096 * 4096);
}

// Large host batches: sub-batches are copied to device staging buffers on a separate stream while the
// chain processes the previous one in place, so the PCIe transfer overlaps the kernels.
static int push_pipelined(ldvb_handle *h, const uint8_t *host, size_t n) {
  const size_t bps = h->s_raw.elem;
  { int rcs = stage_init(h); if (rcs) return rcs; }
  const uint64_t each = sub_batch_size(h, n);
  const uint64_t nsub = (n + each - 1) / each;
  auto issue = [&](uint64_t i) -> int {
    const uint64_t off = i * each, m = std::min<uint64_t>(each, n - off);
    CK(cudaMemcpyAsync(h->d_stage[i & 1].p, host + off * bps, m * bps, cudaMemcpyHostToDevice, h->copy_st));
    CK(cudaEventRecord(h->copy_done[i & 1], h->copy_st));
    return LDVB_OK;
  };
  int rc = issue(0);
  if (rc) return rc;
  for (uint64_t i = 0; i < nsub; ++i) {
    CK(cudaStreamWaitEvent(h->st, h->copy_done[i & 1], 0));
    if (i + 1 < nsub && (rc = issue(i + 1))) return rc;   // the other buffer is idle: run_chain is synchronous
    const uint64_t m = std::min<uint64_t>(each, n - i * each);
    uint64_t got = 0;
    if ((rc = run_chain(h, h->d_stage[i & 1].p, true, m, h->d_ts.as<uint8_t>(), h->ts_cap, &got))) return rc;
    if ((rc = push_collect(h, got))) return rc;
  }
  CK(cudaStreamSynchronize(h->st));
  push_publish(h);
  return LDVB_OK;
}

// Moves the telemetry of the finished work to the published copies.  Caller holds amu; the live copies are
// not in use (the caller is the worker between two jobs, or the worker is idle).
static void telemetry_publish(ldvb_handle *h) {
  h->meas.kernel_launches = h->launches;
  h->meas_pub = h->meas;
  auto move_all = [](std::vector<float> &from, std::vector<float> &to) {
    to.insert(to.end(), from.begin(), from.end());
    from.clear();
  };
  move_all(h->cnr_queue, h->cnr_pub);
  move_all(h->vber_queue, h->vber_pub);
  move_all(h->spec_queue, h->spec_pub);
}

// Telemetry pulls while a worker exists: never wait for the chain; serve what finished jobs published.
static size_t telemetry_take(ldvb_handle *h, std::vector<float> ldvb_handle::*pub, float *dst, size_t cap, size_t unit) {
  std::lock_guard<std::mutex> lk(h->amu);
  if (h->pending == 0) { meas_finish(h); telemetry_publish(h); }
  std::vector<float> &q = h->*pub;
  const size_t k = std::min(cap, q.size() / unit);
  if (k && dst) memcpy(dst, q.data(), k * unit * 4);
  q.erase(q.begin(), q.begin() + k * unit);
  return k;
}

// ---- async_push: the worker thread
static void async_worker(ldvb_handle *h) {
  cudaSetDevice(h->cfg.device);
  for (;;) {
    ldvb_handle::Job job;
    {
      std::unique_lock<std::mutex> lk(h->amu);
      h->acv_job.wait(lk, [&] { return h->worker_stop || !h->jobs.empty(); });
      if (h->jobs.empty()) return;                     // stop requested and nothing left
      job = h->jobs.front();
      h->jobs.pop_front();
    }
    int rc = h->worker_rc;                              // after an error the remaining jobs are dropped
    const auto tj0 = std::chrono::steady_clock::now();
    if (!rc) {
      rc = (cudaStreamWaitEvent(h->st, h->copy_done[job.stage], 0) == cudaSuccess) ? LDVB_OK : LDVB_ECUDA;
      uint64_t got = 0;
      if (!rc) rc = run_chain(h, h->d_stage[job.stage].p, true, job.n, h->d_ts.as<uint8_t>(), h->ts_cap, &got);
      if (!rc) rc = push_collect(h, got);
      // the staging buffer (and the carry copied out of it) and the packets are final once the stream is idle
      if (cudaStreamSynchronize(h->st) != cudaSuccess && !rc) rc = LDVB_ECUDA;
      if (!rc) push_publish(h);
      if (!rc) rc = meas_finish(h);
    }
    {
      std::lock_guard<std::mutex> lk(h->amu);
      if (!rc) telemetry_publish(h);
      if (rc && !h->worker_rc) h->worker_rc = rc;
      h->dbg_chain_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tj0).count();
      ++h->dbg_jobs;
      h->stage_busy[job.stage] = false;
      --h->pending;
    }
    h->acv_done.notify_all();
  }
}

static int push_async(ldvb_handle *h, const uint8_t *host, size_t n) {
  const size_t bps = h->s_raw.elem;
  { int rcs = stage_init(h); if (rcs) return rcs; }
  if (!h->worker_started) {
    h->worker = std::thread(async_worker, h);
    h->worker_started = true;
  }
  const uint64_t each = sub_batch_size(h, n);
  const uint64_t nsub = (n + each - 1) / each;
  int last_stage = -1;
  for (uint64_t i = 0; i < nsub; ++i) {
    const uint64_t off = i * each, m = std::min<uint64_t>(each, n - off);
    const int st = (int)(h->stage_next++ % ldvb_handle::kStages);
    {
      const auto tw0 = std::chrono::steady_clock::now();
      std::unique_lock<std::mutex> lk(h->amu);
      h->acv_done.wait(lk, [&] { return !h->stage_busy[st] || h->worker_rc; });
      if (h->worker_rc) return h->worker_rc;
      h->stage_busy[st] = true;
      h->dbg_stage_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
    }
    CK(cudaMemcpyAsync(h->d_stage[st].p, host + off * bps, m * bps, cudaMemcpyHostToDevice, h->copy_st));
    CK(cudaEventRecord(h->copy_done[st], h->copy_st));
    {
      std::lock_guard<std::mutex> lk(h->amu);
      h->jobs.push_back({st, m});
      ++h->pending;
    }
    h->acv_job.notify_one();
    last_stage = st;
  }
  // The caller may reuse its buffer when we return: wait for the last copy (not for the chain).
  const auto tc0 = std::chrono::steady_clock::now();
  if (last_stage >= 0) CK(cudaEventSynchronize(h->copy_done[last_stage]));
  h->dbg_copy_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count();
  return LDVB_OK;
}

int ldvb_flush(ldvb_handle *h) {
  if (!h) return LDVB_EINVAL;
  const int rc = async_wait(h);
  static const bool dbg = getenv("LDVB_ASYNC_DEBUG") != nullptr;
  if (dbg && h->worker_started)
    fprintf(stderr, "[ldvb async] jobs %llu: chain %.2f ms in the worker, caller waited %.2f ms for a staging buffer, %.2f ms for its last copies\n",
            (unsigned long long)h->dbg_jobs, h->dbg_chain_ms, h->dbg_stage_wait_ms, h->dbg_copy_wait_ms);
  return rc;
}

int ldvb_push(ldvb_handle *h, const void *iq_host, size_t n) {
  if (!h || (!iq_host && n)) return LDVB_EINVAL;
  if (n > h->cfg.max_batch) return fail(h, LDVB_EOVERFLOW, "n_samples exceeds max_batch");
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return fail(h, LDVB_ECUDA, "cudaSetDevice");
  if (h->cfg.async_push && h->sub_batch && n) return push_async(h, static_cast<const uint8_t *>(iq_host), n);
  if (h->sub_batch && n > h->sub_batch + h->sub_batch / 2)
    return push_pipelined(h, static_cast<const uint8_t *>(iq_host), n);
  Stream &raw = h->s_raw;
  if (raw.count + n > raw.cap) return fail(h, LDVB_EOVERFLOW, "raw stream overflow");
  if (n) CK(cudaMemcpyAsync(raw.at(raw.count), iq_host, n * raw.elem, cudaMemcpyHostToDevice, h->st));
  uint64_t got = 0;
  int rc = run_chain(h, nullptr, false, n, h->d_ts.as<uint8_t>(), h->ts_cap, &got);
  if (rc) return rc;
  if ((rc = push_collect(h, got))) return rc;
  CK(cudaStreamSynchronize(h->st));
  push_publish(h);
  return LDVB_OK;
}

int ldvb_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return LDVB_EINVAL;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return LDVB_OK; }
  if (e != cudaSuccess) { cudaGetLastError(); return LDVB_ECUDA; }
  return LDVB_OK;
}

int ldvb_host_unregister(void *ptr) {
  if (!ptr) return LDVB_EINVAL;
  if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return LDVB_ECUDA; }
  return LDVB_OK;
}

int ldvb_pull(ldvb_handle *h, uint8_t *ts_host, size_t cap_packets, size_t *n_packets) {
  if (!h || !n_packets) return LDVB_EINVAL;
  std::lock_guard<std::mutex> lk(h->qmu);
  const size_t avail = (h->ts_queue_ready - h->ts_queue_rd) / 188;
  const size_t n = std::min(avail, cap_packets);
  if (n && ts_host) memcpy(ts_host, h->ts_queue + h->ts_queue_rd, n * 188);
  h->ts_queue_rd += n * 188;
  if (h->ts_queue_rd == h->ts_queue_wr) h->ts_queue_rd = h->ts_queue_wr = h->ts_queue_ready = 0;
  *n_packets = n;
  return LDVB_OK;
}

int ldvb_process_device(ldvb_handle *h, const void *iq_dev, size_t n, uint8_t *ts_dev, size_t cap_packets,
                        size_t *n_packets) {
  if (!h || !n_packets || (!iq_dev && n) || !ts_dev) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  uint64_t got = 0;
  int rc = run_chain(h, iq_dev, true, n, ts_dev, cap_packets, &got);
  *n_packets = (size_t)got;
  return rc;
}

int ldvb_get_meas(ldvb_handle *h, ldvb_meas *m) {
  if (!h || !m) return LDVB_EINVAL;
  if (h->worker_started) {               // async_push: as of the last finished sub-batch, without waiting
    std::lock_guard<std::mutex> lk(h->amu);
    if (h->pending == 0) telemetry_publish(h);
    *m = h->meas_pub;
    return h->worker_rc;
  }
  h->meas.kernel_launches = h->launches;
  *m = h->meas;
  return LDVB_OK;
}

int ldvb_tap(ldvb_handle *h, int which, void *dst, size_t cap, size_t *n_bytes) {
  if (!h || which < 0 || which > 8 || !n_bytes) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  if (!h->cfg.keep_taps) return fail(h, LDVB_ESTATE, "handle created without keep_taps");
  Tap &t = h->taps[which];
  *n_bytes = (size_t)t.bytes;
  if (!dst) return LDVB_OK;
  if (cap < t.bytes) return LDVB_EOVERFLOW;
  if (t.bytes) {
    CK(cudaStreamSynchronize(h->st));
    CK(cudaMemcpy(dst, t.buf.p, t.bytes, cudaMemcpyDeviceToHost));
  }
  return LDVB_OK;
}

static int host_table(const ldvb_config &c, int which, std::vector<uint8_t> &blob) {
  auto put = [&](const void *p, size_t n) { blob.assign((const uint8_t *)p, (const uint8_t *)p + n); };
  // Sample rate seen by the receiver (leandvb.cc:353-399).
  float Fs = c.Fs;
  int decim = 1;
  std::vector<float> fir;
  if (c.resample) { fir = design_resampler(Fs, c.Fm, c.rolloff, c.resample_rej, c.decim, &decim); Fs /= decim; }
  else if (c.decim > 1) Fs /= c.decim;
  switch (which) {
    case LDVB_TABLE_CSTLN: {
      Cstln cs = make_cstln(c.constellation, c.fec, c.hard_metric != 0);
      if (!cs.nsymbols) return LDVB_EINVAL;
      put(cs.cells.data(), cs.cells.size() * sizeof(CstlnCell));
      break;
    }
    case LDVB_TABLE_TRIG16: { auto t = make_trig16(); put(t.data(), t.size() * 4); break; }
    case LDVB_TABLE_RS_EXP: { uint8_t e[512], l[256]; make_rs_tables(e, l); put(e, 512); break; }
    case LDVB_TABLE_RS_LOG: { uint8_t e[512], l[256]; make_rs_tables(e, l); put(l, 256); break; }
    case LDVB_TABLE_DERAND: { auto t = make_derand_pattern(); put(t.data(), t.size()); break; }
    case LDVB_TABLE_FIR: put(fir.data(), fir.size() * 4); break;
    case LDVB_TABLE_FIR_SHIFTED: {
      // The complex taps in force for the first batch: fir_filter::set_freq (dsp.h:270-280) at the frequency
      // run() picks up from the demodulator's initial freq_tap (dsp.h:236-244, leandvb.cc:483-486, 505-510).
      if (!c.resample) return LDVB_EINVAL;
      const float freqw = c.Ftune ? (c.Ftune / Fs) * 65536 : 0.0f;          // set_freq (sdr.h:745-749)
      const float freq_tap = freqw / 65536;
      const float new_freq = freq_tap * (1.0f / decim);
      const float tol = (float)((double)(c.Fm / (Fs * decim)) * 0.1);
      const std::vector<float> sh = shift_taps(fir, fabsf(0.0f - new_freq) > tol ? new_freq : 0.0f);
      put(sh.data(), sh.size() * 4);
      break;
    }
    case LDVB_TABLE_DECONV: {
      DeconvPolys d;
      if (!make_deconv(c.fec, &d)) return LDVB_EINVAL;
      put(d.deconv, 8 * (size_t)d.punctperiod);
      break;
    }
    case LDVB_TABLE_RRC: {
      int steps = 0;
      auto t = design_rrc(Fs, c.Fm, c.rolloff, c.rrc_rej, c.rrc_steps, &steps);
      put(t.data(), t.size() * 4);
      break;
    }
    case LDVB_TABLE_TRELLIS: {
      Trellis t;
      if (!make_trellis(c.fec, &t)) return LDVB_EINVAL;
      for (size_t i = 0; i < t.pred.size(); ++i) { blob.push_back(t.pred[i]); blob.push_back(t.us[i]); }
      break;
    }
    case LDVB_TABLE_VITMAP: {
      Trellis t;
      Cstln cs = make_cstln(c.constellation, c.fec, c.hard_metric != 0);
      if (!cs.nsymbols || !make_trellis(c.fec, &t)) return LDVB_EINVAL;
      VitSyncs v = make_vitsyncs(cs, t);
      blob.push_back((uint8_t)v.nsyncs); blob.push_back((uint8_t)v.nshifts);
      blob.push_back((uint8_t)v.bps); blob.push_back((uint8_t)cs.nsymbols);
      for (int s = 0; s < v.nsyncs; ++s) {
        blob.push_back((uint8_t)v.shift[s]);
        blob.insert(blob.end(), v.map[s].begin(), v.map[s].end());
      }
      break;
    }
    case LDVB_TABLE_HS_POLAR: { const HsTables t = make_hs_tables(); put(t.polar.data(), t.polar.size() * 4); break; }
    case LDVB_TABLE_HS_RECT: { const HsTables t = make_hs_tables(); put(t.rect.data(), t.rect.size() * 2); break; }
    case LDVB_TABLE_HS_SINCOS: { const HsTables t = make_hs_tables(); put(t.sincos.data(), t.sincos.size() * 2); break; }
    default: return LDVB_EINVAL;
  }
  return LDVB_OK;
}

int ldvb_host_table(const ldvb_config *cfg, int which, void *dst, size_t cap, size_t *n_bytes) {
  if (!cfg || !n_bytes) return LDVB_EINVAL;
  std::vector<uint8_t> blob;
  int rc = host_table(*cfg, which, blob);
  if (rc) return rc;
  *n_bytes = blob.size();
  if (!dst) return LDVB_OK;
  if (cap < blob.size()) return LDVB_EOVERFLOW;
  memcpy(dst, blob.data(), blob.size());
  return LDVB_OK;
}

int ldvb_table(ldvb_handle *h, int which, void *dst, size_t cap, size_t *n_bytes) {
  if (!h) return LDVB_EINVAL;
  return ldvb_host_table(&h->cfg, which, dst, cap, n_bytes);
}

int ldvb_get_rx_state(ldvb_handle *h, uint32_t w[22]) {
  if (!h || !w) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  memcpy(w, &h->rx_state, 22 * 4);
  return LDVB_OK;
}

int ldvb_set_rx_state(ldvb_handle *h, const uint32_t w[22]) {
  if (!h || !w) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  memcpy(&h->rx_state, w, 22 * 4);
  h->rx_cold = false;               // the caller supplied a loop state
  return LDVB_OK;
}

namespace {
struct StateBlob {
  uint32_t magic;
  NotchState notch;
  uint32_t rot_index;
  RxState rx;
  HypState hyp[4];
  int32_t locked, skip;
  SyncState sync;
  int32_t derand_pos;
  float fir_current_freq;
};
}  // namespace

size_t ldvb_state_size(const ldvb_handle *) { return sizeof(StateBlob); }

int ldvb_get_state(ldvb_handle *h, void *blob, size_t cap) {
  if (!h || !blob || cap < sizeof(StateBlob)) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  StateBlob b = StateBlob();
  b.magic = 0x4c445642;
  b.notch = h->notch; b.rot_index = h->rot_index; b.rx = h->rx_state;
  for (int i = 0; i < 4; ++i) b.hyp[i] = h->hyp[i];
  b.locked = h->locked; b.skip = h->skip; b.sync = h->sync; b.derand_pos = h->derand_pos;
  b.fir_current_freq = h->fir_current_freq;
  memcpy(blob, &b, sizeof b);
  return LDVB_OK;
}

int ldvb_set_state(ldvb_handle *h, const void *blob, size_t size) {
  if (!h || !blob || size != sizeof(StateBlob)) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  StateBlob b;
  memcpy(&b, blob, sizeof b);
  if (b.magic != 0x4c445642) return LDVB_EINVAL;
  h->notch = b.notch; h->rot_index = b.rot_index; h->rx_state = b.rx; h->rx_cold = false;
  for (int i = 0; i < 4; ++i) h->hyp[i] = b.hyp[i];
  h->locked = b.locked; h->skip = b.skip; h->sync = b.sync; h->derand_pos = b.derand_pos;
  if (h->use_fir && h->fir_current_freq != b.fir_current_freq) {   // the low-pass follows the saved retune
    CK(cudaStreamSynchronize(h->st));
    h->fir_shifted = shift_taps(h->fir_coeffs, b.fir_current_freq);
    h->fir_current_freq = b.fir_current_freq;
    CK(cudaMemcpyAsync(h->d_taps.p, h->fir_shifted.data(), h->fir_shifted.size() * 4, cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
  }
  return LDVB_OK;
}

// ------------------------------------------------------------ CNR / spectrum

int ldvb_pull_cnr(ldvb_handle *h, float *dst, size_t cap, size_t *n) {
  if (!h || !n) return LDVB_EINVAL;
  if (h->worker_started) { *n = telemetry_take(h, &ldvb_handle::cnr_pub, dst, cap, 1); return LDVB_OK; }
  { int rc = meas_finish(h); if (rc) return rc; }
  const size_t k = std::min(cap, h->cnr_queue.size());
  if (k && dst) memcpy(dst, h->cnr_queue.data(), k * 4);
  h->cnr_queue.erase(h->cnr_queue.begin(), h->cnr_queue.begin() + k);
  *n = k;
  return LDVB_OK;
}

int ldvb_pull_vber(ldvb_handle *h, float *dst, size_t cap, size_t *n) {
  if (!h || !n) return LDVB_EINVAL;
  if (h->worker_started) { *n = telemetry_take(h, &ldvb_handle::vber_pub, dst, cap, 1); return LDVB_OK; }
  const size_t k = std::min(cap, h->vber_queue.size());
  if (k && dst) memcpy(dst, h->vber_queue.data(), k * 4);
  h->vber_queue.erase(h->vber_queue.begin(), h->vber_queue.begin() + k);
  *n = k;
  return LDVB_OK;
}

int ldvb_pull_spectrum(ldvb_handle *h, float *dst, size_t cap_rows, size_t *n_rows) {
  if (!h || !n_rows) return LDVB_EINVAL;
  if (h->worker_started) { *n_rows = telemetry_take(h, &ldvb_handle::spec_pub, dst, cap_rows, 1024); return LDVB_OK; }
  { int rc = meas_finish(h); if (rc) return rc; }
  const size_t k = std::min(cap_rows, h->spec_queue.size() / 1024);
  if (k && dst) memcpy(dst, h->spec_queue.data(), k * 1024 * 4);
  h->spec_queue.erase(h->spec_queue.begin(), h->spec_queue.begin() + k * 1024);
  *n_rows = k;
  return LDVB_OK;
}

// -------------------------------------------------------------- time sharding

size_t ldvb_edge_size(void) { return sizeof(EdgeBlob); }

size_t ldvb_shard_min_halo(const ldvb_handle *h) { return h ? (size_t)shard_min_halo(h) : 0; }

int ldvb_shard_detect(ldvb_handle *h, ldvb_shard *s) {
  int rc = shard_check(h, s);
  { int rcw = async_wait(h); if (rcw) return rcw; }
  if (rc) return rc;
  return shard_detect(h, s);
}

int ldvb_shard_front(ldvb_handle *h, const ldvb_shard *s) {
  int rc = shard_check(h, s);
  { int rcw = async_wait(h); if (rcw) return rcw; }
  if (rc) return rc;
  return shard_front(h, s);
}

int ldvb_shard_back(ldvb_handle *h, const void *edge_in, uint8_t *ts_dev, size_t cap_packets, size_t *n_packets,
                    void *edge_out) {
  if (!h || !ts_dev || !n_packets) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return fail(h, LDVB_ECUDA, "cudaSetDevice");
  uint64_t got = 0;
  int rc = shard_back(h, static_cast<const EdgeBlob *>(edge_in), ts_dev, cap_packets, &got, static_cast<EdgeBlob *>(edge_out));
  *n_packets = (size_t)got;
  return rc;
}

// ------------------------------------------------------------ stand-alone stages

// ---------------------------------------------------------------- the ring (NCCL)
// "a single NCCL send/recv over NVLink carrying only the chunk-edge samples and Viterbi/PLL carry
// state".  ldvb_ring_round is one rank's turn: halo and notch bins in (early communicator), detect, halo and bins
// out, the speculative front stage, EDGE in (edge communicator), the exact back stage, EDGE out.  Two
// communicators = two NCCL streams, so that the early messages of the next round travel while an EDGE is still
// awaited.  Messages between two neighbours are matched in issue order and every rank issues them in chunk order,
// all dependencies point forward in the stream (the wrap-around N-1 -> 0 included), so the schedule cannot lock.
// NCCL is resolved at run time (dlopen): a single-GPU user of the library does not need it.
}  // extern "C"

#include <dlfcn.h>

namespace ldvb {
namespace {
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclUint8 = 1;   // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *lib = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return;
    auto sym = [&](const char *n) { return dlsym(lib, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
  });
  return &api;
}
}  // namespace
}  // namespace ldvb

struct RingState {
  void *comm_early = nullptr, *comm_edge = nullptr;
  cudaStream_t st_early = nullptr, st_edge = nullptr;
  int rank = 0, n = 1;
  DevBuf d_bins_in, d_bins_out, d_edge_in, d_edge_out;
  uint8_t *h_edge_in = nullptr, *h_edge_out = nullptr;
  int32_t *h_bins = nullptr;          // [8] pinned: bins in, bins out
  double ms[4] = {0, 0, 0, 0};        // early, front, wait_edge, back (host wall clock, accumulated)
};

namespace {
int ring_fail(ldvb_handle *h, int nccl_rc, const char *what) {
  ldvb::NcclApi *api = ldvb::nccl_api();
  std::string msg = std::string("NCCL ") + what + ": " + (api->GetErrorString ? api->GetErrorString(nccl_rc) : "error");
  return fail(h, LDVB_ECUDA, msg.c_str());
}
#define NK(call, what) do { int rcn_ = (call); if (rcn_ != 0) return ring_fail(h, rcn_, what); } while (0)
}  // namespace

extern "C" {

int ldvb_ring_unique_id(void *id, size_t cap) {
  ldvb::NcclApi *api = ldvb::nccl_api();
  if (!id || cap < sizeof(ldvb::NcclApi::UniqueId)) return LDVB_EINVAL;
  if (!api->ok) return LDVB_ENODEV;
  ldvb::NcclApi::UniqueId u;
  if (api->GetUniqueId(&u) != 0) return LDVB_ECUDA;
  memcpy(id, &u, sizeof u);
  return LDVB_OK;
}

int ldvb_ring_init(ldvb_handle *h, const void *id_early, const void *id_edge, int rank, int nranks) {
  using namespace ldvb;
  if (!h || !id_early || !id_edge || nranks < 1 || rank < 0 || rank >= nranks) return LDVB_EINVAL;
  { int rcw = async_wait(h); if (rcw) return rcw; }
  NcclApi *api = nccl_api();
  if (!api->ok) return fail(h, LDVB_ENODEV, "libnccl.so.2 not found");
  if (h->ring) return fail(h, LDVB_ESTATE, "ring already initialised");
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return fail(h, LDVB_ECUDA, "cudaSetDevice");
  RingState *r = new RingState();
  r->rank = rank; r->n = nranks;
  h->ring = r;
  NcclApi::UniqueId a, b;
  memcpy(&a, id_early, sizeof a); memcpy(&b, id_edge, sizeof b);
  NK(api->CommInitRank(&r->comm_early, nranks, a, rank), "ncclCommInitRank (early)");
  NK(api->CommInitRank(&r->comm_edge, nranks, b, rank), "ncclCommInitRank (edge)");
  CK(cudaStreamCreateWithFlags(&r->st_early, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&r->st_edge, cudaStreamNonBlocking));
  CK(r->d_bins_in.alloc(64)); CK(r->d_bins_out.alloc(64));
  CK(r->d_edge_in.alloc(sizeof(EdgeBlob) + 256)); CK(r->d_edge_out.alloc(sizeof(EdgeBlob) + 256));
  CK(cudaHostAlloc((void **)&r->h_edge_in, sizeof(EdgeBlob), cudaHostAllocDefault));
  CK(cudaHostAlloc((void **)&r->h_edge_out, sizeof(EdgeBlob), cudaHostAllocDefault));
  CK(cudaHostAlloc((void **)&r->h_bins, 64, cudaHostAllocDefault));
  return LDVB_OK;
}

int ldvb_ring_flush(ldvb_handle *h) {
  using namespace ldvb;
  if (!h || !h->ring) return LDVB_EINVAL;
  CK(cudaStreamSynchronize(h->ring->st_early));
  CK(cudaStreamSynchronize(h->ring->st_edge));
  return LDVB_OK;
}

int ldvb_ring_round(ldvb_handle *h, ldvb_shard *s, uint8_t *ts_dev, size_t cap_packets, size_t *n_packets) {
  using namespace ldvb;
  if (!h || !s || !ts_dev || !n_packets) return LDVB_EINVAL;
  RingState *r = h->ring;
  if (!r) return fail(h, LDVB_ESTATE, "ldvb_ring_init first");
  { int rcw = async_wait(h); if (rcw) return rcw; }
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return fail(h, LDVB_ECUDA, "cudaSetDevice");
  NcclApi *api = nccl_api();
  const bool first = (s->n_halo == 0 && s->abs_raw0 == 0), last = s->last != 0;
  const int prv = (r->rank + r->n - 1) % r->n, nxt = (r->rank + 1) % r->n;
  const size_t bps = h->s_raw.elem;
  uint8_t *iq = static_cast<uint8_t *>(const_cast<void *>(s->iq_dev));
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  auto t0 = now();
  // The halo of the previous round has left the caller's buffer (the caller refills it between rounds).
  CK(cudaStreamSynchronize(r->st_early));
  // ---- early: halo + notch bins from the previous chunk
  for (int k = 0; k < 4; ++k) s->bins_before[k] = -1;
  if (!first) {
    NK(api->GroupStart(), "ncclGroupStart");
    NK(api->Recv(iq, s->n_halo * bps, kNcclUint8, prv, r->comm_early, r->st_early), "ncclRecv (halo)");
    NK(api->Recv(r->d_bins_in.p, 16, kNcclUint8, prv, r->comm_early, r->st_early), "ncclRecv (bins)");
    NK(api->GroupEnd(), "ncclGroupEnd");
    CK(cudaMemcpyAsync(r->h_bins, r->d_bins_in.p, 16, cudaMemcpyDeviceToHost, r->st_early));
    CK(cudaStreamSynchronize(r->st_early));
    for (int k = 0; k < 4; ++k) s->bins_before[k] = r->h_bins[k];
  }
  int rc = shard_check(h, s);
  if (rc) return rc;
  if ((rc = shard_detect(h, s))) return rc;
  if (!last) {
    for (int k = 0; k < 4; ++k) r->h_bins[4 + k] = s->bins_after[k];
    CK(cudaMemcpyAsync(r->d_bins_out.p, r->h_bins + 4, 16, cudaMemcpyHostToDevice, r->st_early));
    NK(api->GroupStart(), "ncclGroupStart");
    NK(api->Send(iq + (s->n_halo + s->n_chunk - s->n_halo_next) * bps, s->n_halo_next * bps, kNcclUint8, nxt, r->comm_early, r->st_early), "ncclSend (halo)");
    NK(api->Send(r->d_bins_out.p, 16, kNcclUint8, nxt, r->comm_early, r->st_early), "ncclSend (bins)");
    NK(api->GroupEnd(), "ncclGroupEnd");
  }
  r->ms[0] += ms_since(t0); t0 = now();
  // ---- front (speculative, concurrent on all ranks)
  if ((rc = shard_front(h, s))) return rc;
  r->ms[1] += ms_since(t0); t0 = now();
  // ---- EDGE of the previous chunk
  if (!first) {
    NK(api->Recv(r->d_edge_in.p, sizeof(EdgeBlob), kNcclUint8, prv, r->comm_edge, r->st_edge), "ncclRecv (EDGE)");
    CK(cudaMemcpyAsync(r->h_edge_in, r->d_edge_in.p, sizeof(EdgeBlob), cudaMemcpyDeviceToHost, r->st_edge));
  }
  CK(cudaStreamSynchronize(r->st_edge));      // (also: the previous round's EDGE has left h_edge_out / d_edge_out)
  r->ms[2] += ms_since(t0); t0 = now();
  // ---- back (exact, chained) and the EDGE for the next chunk
  uint64_t got = 0;
  rc = shard_back(h, first ? nullptr : reinterpret_cast<const EdgeBlob *>(r->h_edge_in), ts_dev, cap_packets, &got,
                  last ? nullptr : reinterpret_cast<EdgeBlob *>(r->h_edge_out));
  *n_packets = (size_t)got;
  if (rc) return rc;
  if (!last) {
    CK(cudaMemcpyAsync(r->d_edge_out.p, r->h_edge_out, sizeof(EdgeBlob), cudaMemcpyHostToDevice, r->st_edge));
    NK(api->Send(r->d_edge_out.p, sizeof(EdgeBlob), kNcclUint8, nxt, r->comm_edge, r->st_edge), "ncclSend (EDGE)");
  }
  r->ms[3] += ms_since(t0);
  return LDVB_OK;
}

int ldvb_ring_stats(ldvb_handle *h, double ms4[4], int reset) {
  if (!h || !h->ring || !ms4) return LDVB_EINVAL;
  for (int k = 0; k < 4; ++k) { ms4[k] = h->ring->ms[k]; if (reset) h->ring->ms[k] = 0; }
  return LDVB_OK;
}

int ldvb_ring_destroy(ldvb_handle *h) {
  using namespace ldvb;
  if (!h) return LDVB_EINVAL;
  RingState *r = h->ring;
  if (!r) return LDVB_OK;
  cudaSetDevice(h->cfg.device);
  if (r->st_early) cudaStreamSynchronize(r->st_early);
  if (r->st_edge) cudaStreamSynchronize(r->st_edge);
  NcclApi *api = nccl_api();
  if (r->comm_early) api->CommDestroy(r->comm_early);
  if (r->comm_edge) api->CommDestroy(r->comm_edge);
  if (r->st_early) cudaStreamDestroy(r->st_early);
  if (r->st_edge) cudaStreamDestroy(r->st_edge);
  r->d_bins_in.release(); r->d_bins_out.release(); r->d_edge_in.release(); r->d_edge_out.release();
  if (r->h_edge_in) cudaFreeHost(r->h_edge_in);
  if (r->h_edge_out) cudaFreeHost(r->h_edge_out);
  if (r->h_bins) cudaFreeHost(r->h_bins);
  delete r;
  h->ring = nullptr;
  return LDVB_OK;
}

int ldvb_fir_cf32(int device, const float *x, size_t n_in, const float *taps, uint32_t ntaps, uint32_t decim,
                  float *y, size_t cap_out, size_t *n_out) {
  if (!x || !taps || !y || !n_out || !ntaps || !decim) return LDVB_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return LDVB_ENODEV;
  *n_out = 0;
  if (n_in < ntaps) return LDVB_OK;
  const size_t count = (n_in - ntaps) / decim;
  if (count > cap_out) return LDVB_EOVERFLOW;
  if (!count) return LDVB_OK;
  DevBuf dx, dt, dy;
  int rc = LDVB_OK;
  if (dx.alloc(n_in * 8 + 512) != cudaSuccess || dt.alloc((size_t)ntaps * 8) != cudaSuccess ||
      dy.alloc(count * 8) != cudaSuccess) rc = LDVB_ENOMEM;
  if (!rc) {
    cudaMemset(dx.p, 0, dx.bytes);
    cudaMemcpy(dx.p, x, n_in * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dt.p, taps, (size_t)ntaps * 8, cudaMemcpyHostToDevice);
    FrontendArgs a;
    memset(&a, 0, sizeof a);
    a.src.head = dx.p; a.src.head_count = n_in; a.src.main = nullptr; a.src.c0 = 0;
    a.fmt = 5; a.scale = 1; a.taps = dt.as<float2>(); a.ntaps = ntaps; a.decim = decim;
    a.out = dy.as<float2>(); a.count = count;
    if (launch_frontend(a, 0) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) rc = LDVB_ECUDA;
    else { cudaMemcpy(y, dy.p, count * 8, cudaMemcpyDeviceToHost); *n_out = count; }
  }
  dx.release(); dt.release(); dy.release();
  return rc;
}

static int rs_common(int device, const uint8_t *src, size_t src_bytes, size_t npk, bool deint, uint8_t *rts,
                     int32_t *flags) {
  if (cudaSetDevice(device) != cudaSuccess) return LDVB_ENODEV;
  if (!npk) return LDVB_OK;
  uint8_t gexp[512], glog[256];
  make_rs_tables(gexp, glog);
  DevBuf ds, de, dl, dr, df;
  int rc = LDVB_OK;
  if (ds.alloc(src_bytes + 256) != cudaSuccess || de.alloc(512) != cudaSuccess || dl.alloc(256) != cudaSuccess ||
      dr.alloc(npk * 188) != cudaSuccess || df.alloc(npk * 8) != cudaSuccess) rc = LDVB_ENOMEM;
  if (!rc) {
    cudaMemcpy(ds.p, src, src_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(de.p, gexp, 512, cudaMemcpyHostToDevice);
    cudaMemcpy(dl.p, glog, 256, cudaMemcpyHostToDevice);
    cudaError_t e;
    if (deint) {
      DeintRsArgs a;
      a.mpeg = ds.as<uint8_t>(); a.npackets = npk; a.gf_exp = de.as<uint8_t>(); a.gf_log = dl.as<uint8_t>();
      a.rs_out = nullptr; a.rts_out = dr.as<uint8_t>(); a.flags = df.as<int32_t>();
      e = launch_deint_rs(a, 0);
    } else {
      RsOnlyArgs a;
      a.rs_in = ds.as<uint8_t>(); a.npackets = npk; a.gf_exp = de.as<uint8_t>(); a.gf_log = dl.as<uint8_t>();
      a.rts_out = dr.as<uint8_t>(); a.flags = df.as<int32_t>();
      e = launch_rs_only(a, 0);
    }
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) rc = LDVB_ECUDA;
    else {
      cudaMemcpy(rts, dr.p, npk * 188, cudaMemcpyDeviceToHost);
      if (flags) cudaMemcpy(flags, df.p, npk * 8, cudaMemcpyDeviceToHost);
    }
  }
  ds.release(); de.release(); dl.release(); dr.release(); df.release();
  return rc;
}

int ldvb_deint_rs(int device, const uint8_t *mpeg, size_t n_bytes, uint8_t *rts, size_t cap_packets, size_t *n_packets,
                  int32_t *flags) {
  if (!mpeg || !rts || !n_packets) return LDVB_EINVAL;
  size_t npk = (n_bytes >= 2448) ? (n_bytes - 2244) / 204 : 0;
  if (npk > cap_packets) return LDVB_EOVERFLOW;
  *n_packets = npk;
  return rs_common(device, mpeg, n_bytes, npk, true, rts, flags);
}

int ldvb_rs_decode(int device, const uint8_t *rs204, size_t npk, uint8_t *ts188, int32_t *flags) {
  if (!rs204 || !ts188) return LDVB_EINVAL;
  return rs_common(device, rs204, npk * 204, npk, false, ts188, flags);
}

}  // extern "C"
