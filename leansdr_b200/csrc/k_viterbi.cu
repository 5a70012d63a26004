// k_viterbi.cu -- K5: viterbi_sync (dvb.h:1173-1416) over the explicit trellis of
// viterbi.h:43-293, all code rates, hypothesis tracking included.
//
// One warp = one decoder (sync hypothesis); the 64 states are spread two per
// lane.  Every FEC block the warp performs the reference's add-compare-select
// with the SAME candidate order and tie rules:
//   1. the branch whose label equals the received coded symbol, metric
//      cost[pred] + cost (viterbi.h:207-219, `<=` against max);
//   2. every existing branch in increasing label order with metric cost[pred]
//      (the "rescan", viterbi.h:221-234): `<=`, so later candidates win ties;
// then best / second-best state (first minimum wins, duplicates count for the
// second best: viterbi.h:239-245), normalisation by the minimum, quality =
// second - best, output = oldest symbol of the best state's path register
// (register-exchange survivor memory, viterbi.h:283-289).
// The current decoder runs on every 128-block chunk, the others only on every
// `resync_period`-th chunk with the state they had 32 chunks earlier; the best
// sum of quality over blocks >= discr_delay becomes current (dvb.h:1386-1411).
//
// Rescan without the dead labels.  All rescan candidates of one predecessor carry
// the same metric, and among equal metrics the LAST label wins: per state only
// (pred, us) of the largest label of every distinct predecessor matters, in
// increasing order of that label.  The CTA builds these lists from the full
// trellis when it starts (7/8: 64 entries per state instead of 256 labels); the
// selected branch is the one the reference selects, tie for tie.
//
// Time parallelism.  The recurrence is serial, but a Viterbi decoder forgets: once all
// survivors descend from one state, the normalised metrics and the path registers no longer
// depend on where the decoder started.  The stream is cut into segments (whole re-sync
// groups); one CTA per segment starts its decoders COLD a little earlier -- all of them on the
// previous `warm_others` re-sync chunks, which is the only data the non-current ones see
// between two votes anyway, with those chunks' votes; then the current one again, cold, on the
// last `warm_chunks` chunks -- and records the state with which it enters its segment.  k_vit_verify compares entry(g) with exit(g-1) bit for bit (64 metrics, 64 path
// registers per decoder, the current decoder and the re-sync phase); a segment that did not
// merge is re-run from its predecessor's exit state (list mode).  Exactness therefore holds by
// induction, as for the notch segments and the receiver spans; segment 0 always starts from
// the carried state.
//
// Instances (device code in k_vit_dev.cuh, which the host shim of tests/emu compiles too):
//   k_viterbi<kVitGeneric>  any trellis: CTA per segment, warp per hypothesis, rescan lists;
//   k_viterbi<kVitFull>     every state a predecessor of every state (7/8): no rescan walk at all, its outcome is the
//                           minimum of the current metrics plus the tie rule (2.3 GS/s against 0.24 on B200);
//   k_viterbi<kVitR12>      rate 1/2, trellis rows in registers (LDVB_VIT_WS=0);
//   k_viterbi_ws            rate 1/2, one WARP per segment, its decoders side by side in shared memory: every resident
//                           warp walks a decoder (4.2 GS/s against 2.1).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ldvb {

namespace {

#include "k_vit_dev.cuh"

}  // namespace

int vit_rescan_entries(int bits_in) { return bits_in >= 6 ? 64 : (1 << bits_in); }

static size_t vit_smem_bytes(int ncs, int nb, int nsyncs) {
  return (((size_t)128 * ncs + (size_t)128 * nb + 15) & ~(size_t)15) +
         (size_t)nsyncs * (2 * 64 * 4 + 2 * 64 * 8 + 4 + kVitChunk * 5) + 96;
}

static cudaError_t vit_configure(size_t smem) {
  static PerDeviceMark configured;   // per device (function attributes belong to the context)
  if (smem > 48 * 1024 && configured.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(k_viterbi<kVitGeneric>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_viterbi<kVitFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.commit(smem);
  }
  return cudaSuccess;
}

// Every state a predecessor of every state (the precondition of k_viterbi<kVitFull>; 7/8 in DVB-S): checked on the
// host trellis, not assumed from the code rate.  LDVB_VIT_FULL=0 keeps the 64-entry rescan walk.
bool vit_trellis_is_full(const uint8_t *pred, int ncs, int bits_in) {
  static const bool enabled = [] { const char *e = getenv("LDVB_VIT_FULL"); return !(e && e[0] == '0'); }();
  if (!enabled || vit_rescan_entries(bits_in) != 64 || ncs < 64) return false;
  for (int s = 0; s < 64; ++s) {
    unsigned long long seen = 0;
    for (int c = 0; c < ncs; ++c) { const int p = pred[s * ncs + c]; if (p < 64) seen |= 1ull << p; }
    if (seen != ~0ull) return false;
  }
  return true;
}

// Rate 1/2 (4 labels, 2 predecessors per state, at most 4 decoders) runs one warp per segment (k_viterbi_ws);
// LDVB_VIT_WS=0 keeps one CTA per segment.
static bool vit_use_ws(int ncs, int nb, int nsyncs) {
  static const bool enabled = [] { const char *e = getenv("LDVB_VIT_WS"); return !(e && e[0] == '0'); }();
  return enabled && ncs == 4 && nb == 2 && nsyncs <= 4;
}
static size_t vit_ws_smem_bytes(int ncs, int nb, int nsyncs) {
  return (((size_t)128 * ncs + (size_t)128 * nb + 15) & ~(size_t)15) +
         (size_t)kVitWsWarps * ((size_t)nsyncs * (2 * 64 * 4 + 2 * 64 * 8) + kVitChunk * 4 + 16 * 4 + kVitChunk);
}
static cudaError_t vit_ws_configure() {
  static PerDeviceMark configured;
  if (configured.need(1)) {
    cudaError_t e = cudaFuncSetAttribute(k_viterbi_ws, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    configured.commit(1);
  }
  return cudaSuccess;
}

// Segments that are resident at the same time on the current device: one full wave (CTAs, or warps of k_viterbi_ws).
int vit_resident_segments(int ncs, int bits_in, int nsyncs) {
  int dev = 0, sms = 0, per_sm = 0;
  if (vit_use_ws(ncs, vit_rescan_entries(bits_in), nsyncs)) {
    if (vit_ws_configure() != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_viterbi_ws, 32 * kVitWsWarps,
                                                      vit_ws_smem_bytes(ncs, vit_rescan_entries(bits_in), nsyncs)) != cudaSuccess)
      return 0;
    return sms * per_sm * kVitWsWarps;
  }
  const size_t smem = vit_smem_bytes(ncs, vit_rescan_entries(bits_in), nsyncs);
  if (vit_configure(smem) != cudaSuccess) return 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      (ncs == 4 && nsyncs <= 4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_viterbi<kVitR12>, 32 * nsyncs, smem)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_viterbi<kVitGeneric>, 32 * nsyncs, smem)) != cudaSuccess)
    return 0;
  return sms * per_sm;
}

cudaError_t launch_viterbi(const VitArgs &a, const VitSegArgs &sg, uint32_t nblocks, cudaStream_t st) {
  if (!a.nchunks || !nblocks) return cudaSuccess;
  if (vit_use_ws(a.ncs, sg.nb, a.nsyncs)) {
    cudaError_t e = vit_ws_configure();
    if (e != cudaSuccess) return e;
    k_viterbi_ws<<<(nblocks + kVitWsWarps - 1) / kVitWsWarps, 32 * kVitWsWarps, vit_ws_smem_bytes(a.ncs, sg.nb, a.nsyncs), st>>>(a, sg, nblocks);
    return cudaGetLastError();
  }
  const size_t smem = vit_smem_bytes(a.ncs, sg.nb, a.nsyncs);
  cudaError_t e = vit_configure(smem);
  if (e != cudaSuccess) return e;
  // rate 1/2 (4 labels, 2 predecessors per state) has its trellis rows in registers
  if (a.ncs == 4 && sg.nb == 2 && a.nsyncs <= 4) k_viterbi<kVitR12><<<nblocks, 32 * a.nsyncs, smem, st>>>(a, sg);
  else if (sg.full && sg.nb == 64) k_viterbi<kVitFull><<<nblocks, 32 * a.nsyncs, smem, st>>>(a, sg);   // (same launch bounds and
  else k_viterbi<kVitGeneric><<<nblocks, 32 * a.nsyncs, smem, st>>>(a, sg);                           //  shared memory as the generic one)
  return cudaGetLastError();
}

cudaError_t launch_vit_verify(const VitSegArgs &sg, int nsyncs, uint8_t *ok, uint32_t *nfail, cudaStream_t st) {
  if (sg.nseg < 2) return cudaSuccess;
  k_vit_verify<<<sg.nseg - 1, 64, 0, st>>>(sg, nsyncs, ok, nfail);
  return cudaGetLastError();
}

cudaError_t launch_vit_commit(const VitArgs &a, const VitSegArgs &sg, cudaStream_t st) {
  k_vit_commit<<<1, 256, 0, st>>>(a, sg);
  return cudaGetLastError();
}

}  // namespace ldvb
