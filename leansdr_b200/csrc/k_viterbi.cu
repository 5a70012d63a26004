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
// This kernel follows the reference order exactly (serial in time, parallel
// over states and hypotheses), so the output is bit-identical.
#include "common.cuh"
#include "kernels.h"

namespace ldvb {

namespace {

constexpr int kVitChunk = 128;

__global__ void __launch_bounds__(512)
k_viterbi(VitArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  // Layout: trellis pred[64*ncs], us[64*ncs]; per warp: cost[2][64] int32, path[2][64] u64.
  uint8_t *t_pred = smem;
  uint8_t *t_us = t_pred + 64 * a.ncs;
  size_t off = ((size_t)128 * a.ncs + 15) & ~(size_t)15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = a.nsyncs;
  int32_t *cost_all = reinterpret_cast<int32_t *>(smem + off);
  off += (size_t)nw * 2 * 64 * 4;
  uint64_t *path_all = reinterpret_cast<uint64_t *>(smem + off);
  off += (size_t)nw * 2 * 64 * 8;
  int32_t *totaldiscr = reinterpret_cast<int32_t *>(smem + off);
  off += (size_t)nw * 4;
  int *s_ctl = reinterpret_cast<int *>(smem + off);   // [0] current_sync, [1] resync_phase

  for (int i = threadIdx.x; i < 64 * a.ncs; i += blockDim.x) { t_pred[i] = a.trellis_pred[i]; t_us[i] = a.trellis_us[i]; }
  int32_t *cost = cost_all + (size_t)warp * 128;
  uint64_t *path = path_all + (size_t)warp * 128;
  VitDecState *st = a.state + warp;
  int bank = st->bank;
  for (int s = lane; s < 64; s += 32) { cost[bank * 64 + s] = st->cost[s]; path[bank * 64 + s] = st->path[s]; }
  if (threadIdx.x == 0) { s_ctl[0] = a.ctl->current_sync; s_ctl[1] = a.ctl->resync_phase; }
  __syncthreads();

  const uint8_t *map = a.maps + (size_t)warp * a.nsymbols;
  const int shift = a.shifts[warp];
  const int discr_delay = 64 / a.bits_in;   // dvb.h:1369
  const uint64_t path_mask = (1ull << a.path_nbits) - 1;
  const int read_shift = (a.path_depth - 1) * a.path_nbits;
  const int bytes_per_chunk = kVitChunk * a.bits_in / 8;

  for (uint64_t chunk = 0; chunk < a.nchunks; ++chunk) {
    const int current = s_ctl[0];
    const bool resync = (s_ctl[1] == 0);
    const bool mine = (warp == current);
    if (mine || resync) {
      int32_t td = 0;
      uint64_t outstream = 0; int nout = 0;
      uint8_t *outp = a.out + chunk * bytes_per_chunk;
      const uint32_t *pin = a.symbols + chunk * (uint64_t)kVitChunk * a.nshifts + shift;
      for (int blk = 0; blk < kVitChunk; ++blk, pin += a.nshifts) {
        // update_sync (dvb.h:1353-1364): coded symbol and cost of this FEC block
        unsigned cs = 0; int32_t bcost = 0;
        for (int i = 0; i < a.nshifts; ++i) {
          const uint32_t w = pin[i];
          cs = ((cs << a.bps) | map[(w >> 16) & 0xffu]) & 0xffu;
          bcost += (int32_t)(int16_t)(w & 0xffffu);
        }
        const int32_t *cc = cost + bank * 64;
        const uint64_t *pc = path + bank * 64;
        int32_t *cn = cost + (bank ^ 1) * 64;
        uint64_t *pn = path + (bank ^ 1) * 64;
        int32_t my_m[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = lane + 32 * h;
          const uint8_t *row_p = t_pred + s * a.ncs, *row_u = t_us + s * a.ncs;
          int32_t best_m = 0x7fffffff; int best_pred = 0, best_us = 0;
          {
            const int p = row_p[cs];
            if (p != 65) {
              const int32_t m = cc[p] + bcost;
              if (m <= best_m) { best_m = m; best_pred = p; best_us = row_u[cs]; }
            }
          }
          if (a.ncs != 1) {
            for (int c = 0; c < a.ncs; ++c) {
              const int p = row_p[c];
              if (p == 65) continue;
              const int32_t m = cc[p];
              if (m <= best_m) { best_m = m; best_pred = p; best_us = row_u[c]; }
            }
          }
          uint64_t np = pc[best_pred];
          if (a.path32) np = (uint64_t)(uint32_t)(((uint32_t)np << a.path_nbits) | (uint32_t)best_us);
          else np = (np << a.path_nbits) | (uint64_t)best_us;
          pn[s] = np; cn[s] = best_m; my_m[h] = best_m;
        }
        // best state: minimum, first index wins (viterbi.h:239-243)
        int32_t bm; int bs;
        if (my_m[1] < my_m[0]) { bm = my_m[1]; bs = lane + 32; } else { bm = my_m[0]; bs = lane; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const int32_t om = __shfl_xor_sync(0xffffffffu, bm, o);
          const int os = __shfl_xor_sync(0xffffffffu, bs, o);
          if (om < bm || (om == bm && os < bs)) { bm = om; bs = os; }
        }
        // second best: minimum over all states except the best one (duplicates count)
        int32_t b2 = 0x7fffffff;
#pragma unroll
        for (int h = 0; h < 2; ++h) if (lane + 32 * h != bs && my_m[h] < b2) b2 = my_m[h];
#pragma unroll
        for (int o = 16; o; o >>= 1) { const int32_t v = __shfl_xor_sync(0xffffffffu, b2, o); if (v < b2) b2 = v; }
        __syncwarp();
        bank ^= 1;
        // normalise (viterbi.h:249)
        cn[lane] -= bm; cn[lane + 32] -= bm;
        __syncwarp();
        const int32_t quality = b2 - bm;
        if (blk >= discr_delay) td += quality;
        if (mine) {
          const unsigned result = (unsigned)((pn[bs] >> read_shift) & path_mask);
          outstream = (outstream << a.bits_in) | result;
          nout += a.bits_in;
          while (nout >= 8) {
            if (lane == 0) *outp = (uint8_t)(outstream >> (nout - 8));
            ++outp; nout -= 8;
          }
        }
      }
      if (lane == 0) totaldiscr[warp] = td;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (resync) {   // dvb.h:1402-1411
        int best = current;
        for (int s = 0; s < a.nsyncs; ++s) if (totaldiscr[s] > totaldiscr[best]) best = s;
        s_ctl[0] = best;
      }
      if (++s_ctl[1] >= a.resync_period) s_ctl[1] = 0;
    }
    __syncthreads();
  }
  for (int s = lane; s < 64; s += 32) { st->cost[s] = cost[bank * 64 + s]; st->path[s] = path[bank * 64 + s]; }
  if (lane == 0) st->bank = 0;
  // bank is re-based to 0 on store
  if (threadIdx.x == 0) { a.ctl->current_sync = s_ctl[0]; a.ctl->resync_phase = s_ctl[1]; }
}

}  // namespace

cudaError_t launch_viterbi(const VitArgs &a, cudaStream_t st) {
  if (!a.nchunks) return cudaSuccess;
  size_t smem = (((size_t)128 * a.ncs + 15) & ~(size_t)15) + (size_t)a.nsyncs * (2 * 64 * 4 + 2 * 64 * 8 + 4) + 64;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_viterbi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  k_viterbi<<<1, 32 * a.nsyncs, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace ldvb
