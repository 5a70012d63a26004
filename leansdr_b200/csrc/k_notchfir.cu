// k_notchfir.cu -- auto_notch::process (sdr.h:119-138) and the fir_filter that follows it in leandvb's
// graph (leandvb.cc:296-306 -> :353-384, dsp.h:246-259), in ONE pass: p_notched never exists in HBM.
//
// What is serial in auto_notch is only the one-pole chain  estim = bb*k + estim*(1-k)  (two dependent
// operations per sample and component); the products in front of it (bb*k) and everything behind it
// (out = x - estim*e, the FIR) are data parallel.  k_notch_apply (k_notch.cu) lets every lane do all of it for
// its own segment, so the whole kernel runs at the speed of 1-2 warps per scheduler.  Here the roles are split
// inside a CTA of 32 segments ("rows"):
//
//   worker warps (8 x 4 rows, lane = sample of a 32-sample tile):
//     LOAD(t+2)  the rows' next raw tile by 16-byte cp.async (each warp loads the rows it works on)
//     A(t)       convert, bk = (x * conj(e)) * k                        -> shared tile bk[t & 1]
//     CD(t-2)    out = gain * (x - sum estim*e); FIR over the row        <- shared tile est[t & 1]
//   chain warp (lane = row):
//     B(t-1)     estim = bk + estim*(1-k), 32 steps per tile, both tiles as conflict-free float4 columns
//
// one __syncthreads per tile step.  Same arithmetic, same order, same rounding as the reference; segments,
// warm-up from the weighted-sum guess, bit-for-bit verification of entry(j) == exit(j-1) and the repair path
// are those of k_notch.cu (the host code is shared).
//
// FIR on the store path (decimation 1, N <= kFirFuseMaxTaps taps): with u = [carried notched samples | notched
// samples of this batch], y[k] = sum_i taps[i] * u[k + N - i] (dsp.h:250-256), i.e. the output whose NEWEST
// input is batch sample g is y[carry + g - N] = sum_i taps[i] * v[g - i].  A row produces the outputs whose
// N inputs all lie in its own segment; the N-1 outputs that straddle a segment boundary are computed by
// k_fir_edges from the first N-1 / last N notched samples every segment also leaves in `edge` (which is what
// a repaired segment rewrites, too), and the last N samples of the batch become the carry of the next one.
#include "common.cuh"
#include <cstdlib>

#include "kernels.h"
#include "notch_common.cuh"

namespace ldvb {

namespace {

constexpr int kFT = 32;                         // samples per tile = lanes along a row
constexpr int kFRows = 32;                      // segments per CTA
constexpr int kFWorkers = 8;                    // worker warps, kFRows / kFWorkers rows each
constexpr int kFRowsPerWarp = kFRows / kFWorkers;
constexpr int kFThreads = (kFWorkers + 1) * 32; // + the chain warp (warp 0)
constexpr int kFRawStages = 5;                  // raw tiles t+2 (in flight) ... t-2 (read again by CD)
constexpr int kFPitch = (kFT + 2) * 8;          // cf32 tiles: 272 B rows, 16 (mod 128): float4 columns are conflict-free
constexpr int kFTilesPerBlock = kNotchN / kFT;
static_assert(kFirFuseMaxTaps <= kFT, "the FIR history of a tile must fit in the previous tile");

struct FRow {
  int64_t base;                                 // block of local tile 0
  uint64_t own_begin, own_end, run_begin;       // blocks
  uint32_t seg;
  int have, ep0, pad;
};

template <int FMT> struct FRaw {
  static constexpr uint32_t bps = (FMT <= 1) ? 2u : (FMT <= 3 ? 4u : 8u);
  static constexpr int pitch = kFT * (int)bps + 16;     // + one 16-byte piece when the row does not start on a 16-byte boundary
};

template <int FMT, int NSLOTS, bool FIR>
struct FSmem {
  static constexpr size_t raw = (size_t)kFRawStages * kFRows * FRaw<FMT>::pitch;
  static constexpr size_t tile = (size_t)NSLOTS * kFRows * kFPitch;    // one bk / est tile (all slots)
  static constexpr size_t nt = FIR ? (size_t)2 * kFRows * kFPitch : 0; // notched tiles: current + previous
  static constexpr size_t off_bk = raw, off_est = raw + 2 * tile, off_nt = raw + 4 * tile;
  static constexpr size_t off_row = off_nt + nt;
  static constexpr size_t off_taps = off_row + kFRows * sizeof(FRow);
  static constexpr size_t total = off_taps + (size_t)kFirFuseMaxTaps * 8;
};

template <int FMT, int NSLOTS, bool FIR>
__global__ void __launch_bounds__(kFThreads, NSLOTS == 1 ? 2 : 1)
k_notch_fir(NotchFirArgs fa, const uint32_t *seg_list, uint32_t nlist, const float2 *guess) {
  using SM = FSmem<FMT, NSLOTS, FIR>;
  constexpr uint32_t bps = FRaw<FMT>::bps;
  constexpr int kRawPitch = FRaw<FMT>::pitch;
  constexpr uint32_t align_elems = 16 / bps;
  const NotchApplyArgs &a = fa.n;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *s_raw = smem;
  unsigned char *s_bk = smem + SM::off_bk;
  unsigned char *s_est = smem + SM::off_est;
  unsigned char *s_nt = smem + SM::off_nt;
  FRow *s_row = reinterpret_cast<FRow *>(smem + SM::off_row);
  float2 *s_taps = reinterpret_cast<float2 *>(smem + SM::off_taps);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool repair = (seg_list != nullptr);

  // ---- chain warp: segment plans (lane = row) and start states
  float er[NSLOTS], ei[NSLOTS];
#pragma unroll
  for (int s = 0; s < NSLOTS; ++s) { er[s] = 0.f; ei[s] = 0.f; }
  bool forced = false;
  int start_kind = 0;
  if (warp == 0) {
    const uint32_t g = blockIdx.x * kFRows + lane;
    uint32_t seg; bool have;
    if (repair) { have = g < nlist; seg = have ? seg_list[g] : 0; }
    else { seg = g; have = seg < a.nsegs; }
    FRow r;
    r.base = 0; r.own_begin = r.own_end = r.run_begin = 0; r.seg = seg; r.have = have ? 1 : 0; r.ep0 = 0; r.pad = 0;
    if (have) {
      SegPlan p = plan_segment(a, seg);
      start_kind = p.start_kind;
      if (repair) {
        // re-run exactly from the exit state of the predecessor
        const float2 *fe = a.seg_exit + (size_t)(seg - 1) * kNotchMaxSlots;
        forced = true;
        p.run_begin = p.own_begin;
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) { er[s] = fe[s].x; ei[s] = fe[s].y; }
      } else if (p.start_kind == 0) {
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) { er[s] = a.state_in->slot[s].est_re; ei[s] = a.state_in->slot[s].est_im; }
      } else if (p.start_kind == 2) {
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) { const float2 gs = guess_from_sums(a, p, guess, s); er[s] = gs.x; ei[s] = gs.y; }
      }
      r.own_begin = p.own_begin; r.own_end = p.own_end; r.run_begin = p.run_begin; r.ep0 = p.epoch;
      r.base = repair ? (int64_t)p.run_begin
                      : (int64_t)(a.block0 + (uint64_t)seg * a.seg_blocks) - (int64_t)a.warm_blocks;
    }
    s_row[lane] = r;
  }
  if (FIR) for (int i = threadIdx.x; i < fa.fir_n; i += kFThreads) s_taps[i] = fa.taps[i];
  __syncthreads();

  const uint64_t iters = repair ? (uint64_t)a.seg_blocks : (uint64_t)a.warm_blocks + a.seg_blocks;
  const int64_t total = (int64_t)(iters * kFTilesPerBlock);
  const float k = a.k, omk = fsub(1.0f, a.k), gain = a.gain;
  const bool unit_gain = (gain == 1.0f);
  // Samples of the user's buffer (second part of the stream) start at element c0; rows there are fetched from
  // the 16-byte boundary below them: `lead_main` elements in front.
  const uint32_t lead_main = a.src.main ? (uint32_t)((0 - a.src.c0) & (uint64_t)(align_elems - 1)) : 0u;

  if (warp == 0) {
    // ================================================================ chain warp
    const FRow r = s_row[lane];
    int ep = r.ep0;
    for (int64_t s = 0; s < total + 2; ++s) {
      const int64_t t = s - 1;
      if (t >= 0 && t < total) {
        const int64_t blk = r.base + t / kFTilesPerBlock;
        const int tib = (int)(t % kFTilesPerBlock);
        const bool active = r.have && blk >= (int64_t)r.run_begin && blk < (int64_t)r.own_end;
        if (active) {
          if (tib == 0) {
            // Block start: entry snapshot, epoch switch / resets (sdr.h:97-109).
            while (ep + 1 < a.nepochs && a.epochs[ep + 1].first_block <= (uint64_t)blk) ++ep;
            if ((uint64_t)blk == r.own_begin && a.seg_entry && !forced) {
#pragma unroll
              for (int sl = 0; sl < NSLOTS; ++sl) a.seg_entry[(size_t)r.seg * kNotchMaxSlots + sl] = make_float2(er[sl], ei[sl]);
            }
            if (a.epochs[ep].first_block == (uint64_t)blk) {
#pragma unroll
              for (int sl = 0; sl < NSLOTS; ++sl)
                if (a.epochs[ep].reset[sl]) { er[sl] = 0.f; ei[sl] = 0.f; }
            }
          }
          const unsigned char *bk = s_bk + (size_t)(t & 1) * SM::tile + (size_t)lane * kFPitch;
          unsigned char *es = s_est + (size_t)(t & 1) * SM::tile + (size_t)lane * kFPitch;
#pragma unroll
          for (int n = 0; n < kFT; n += 2) {
#pragma unroll
            for (int sl = 0; sl < NSLOTS; ++sl) {
              const float4 b = *reinterpret_cast<const float4 *>(bk + (size_t)sl * kFRows * kFPitch + n * 8);
              const float r0 = fadd(b.x, fmul(er[sl], omk)), i0 = fadd(b.y, fmul(ei[sl], omk));
              er[sl] = fadd(b.z, fmul(r0, omk)); ei[sl] = fadd(b.w, fmul(i0, omk));
              *reinterpret_cast<float4 *>(es + (size_t)sl * kFRows * kFPitch + n * 8) = make_float4(r0, i0, er[sl], ei[sl]);
            }
          }
        }
      }
      __syncthreads();
    }
    if (r.have) {
      if (a.seg_exit) {
#pragma unroll
        for (int sl = 0; sl < NSLOTS; ++sl) a.seg_exit[(size_t)r.seg * kNotchMaxSlots + sl] = make_float2(er[sl], ei[sl]);
      }
      if (a.seg_exact && !forced) a.seg_exact[r.seg] = (start_kind != 2) ? 1 : 0;
    }
    return;
  }

  // ================================================================== worker warps
  // Everything that depends on (row, block) only -- is the row active there, where do its samples lie, which
  // tables are in force, where do its outputs go -- is worked out once per 4096-sample block (128 tiles) and kept
  // in registers; a tile step then costs a handful of instructions per row besides the arithmetic.
  const int w = warp - 1;
  const int row0 = w * kFRowsPerWarp;
  const int N = FIR ? fa.fir_n : 0, L = N > 0 ? N - 1 : 0;
  constexpr int kPieces = kRawPitch / 16;                             // 16-byte pieces of a staged row, incl. the extra one
  constexpr int kLoadIters = (kFRowsPerWarp * kPieces + 31) / 32;
  constexpr uint32_t kStageBytes = (uint32_t)kFRows * kRawPitch;

  // row geometry in local block indices (tile t lies in local block t / 128)
  int lb_run[kFRowsPerWarp], lb_own[kFRowsPerWarp], lb_end[kFRowsPerWarp];
#pragma unroll
  for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
    const FRow &r = s_row[row0 + rr];
    lb_run[rr] = r.have ? (int)((int64_t)r.run_begin - r.base) : 0;
    lb_own[rr] = r.have ? (int)((int64_t)r.own_begin - r.base) : 0;
    lb_end[rr] = r.have ? (int)((int64_t)r.own_end - r.base) : 0;
  }
  // Element offset in front of the rows of a block (rows in the user's buffer start `lead_main` elements after a
  // 16-byte boundary) and the byte address of the block's first sample, aligned down.
  auto block_src = [&](int rr, int lb, uint32_t &lead) -> const unsigned char * {
    uint64_t idx = (uint64_t)(s_row[row0 + rr].base + lb) * kNotchN;
    const unsigned char *part = static_cast<const unsigned char *>(a.src.head);
    lead = 0;
    if (a.src.main && idx >= a.src.c0) { part = static_cast<const unsigned char *>(a.src.main); idx -= a.src.c0; lead = lead_main; }
    return part + (idx - lead) * bps;
  };

  // ---- LOAD state: this lane's pieces of the warp's four rows
  const unsigned char *ld_src[kLoadIters];
  uint32_t ld_dst[kLoadIters];
  uint32_t ld_ok = 0;
  auto load_block = [&](int lb) {
    ld_ok = 0;
#pragma unroll
    for (int i = 0; i < kLoadIters; ++i) {
      const int piece = i * 32 + lane;
      const int rr = piece / kPieces, q = piece % kPieces;
      ld_src[i] = nullptr; ld_dst[i] = 0;
      if (rr < kFRowsPerWarp) {
        // (rr is a per-lane value here: plain shared-memory reads instead of the register copies)
        const FRow &r = s_row[row0 + rr];
        const int64_t blk = r.base + lb;
        if (r.have && blk >= (int64_t)r.run_begin && blk < (int64_t)r.own_end) {
          uint32_t lead;
          const unsigned char *src = block_src(rr, lb, lead);
          if (q < kPieces - 1 || lead) {
            ld_src[i] = src + q * 16;
            ld_dst[i] = (uint32_t)(row0 + rr) * kRawPitch + q * 16;
            ld_ok |= 1u << i;
          }
        }
      }
    }
  };
  uint32_t stL = 0;                                                   // raw stage of the tile being loaded
  auto load = [&](int64_t t) {
    if (t < total) {
      const int tib = (int)(t & (kFTilesPerBlock - 1));
      if (tib == 0) load_block((int)(t >> 7));
      unsigned char *stage = s_raw + stL * kStageBytes;
      const uint32_t toff = (uint32_t)tib * kFT * bps;
#pragma unroll
      for (int i = 0; i < kLoadIters; ++i)
        if ((ld_ok >> i) & 1u) cp_async16(stage + ld_dst[i], ld_src[i] + toff);
    }
    stL = (stL + 1 == kFRawStages) ? 0 : stL + 1;
    cp_async_commit();
  };
  static_assert(kFTilesPerBlock == 128, "t >> 7");

  // ---- A state
  uint32_t actA = 0, xoffA[kFRowsPerWarp];
  const float2 *tabA[kFRowsPerWarp][NSLOTS];
  int epA[kFRowsPerWarp];
  // ---- CD state
  uint32_t actC = 0, firstC = 0, lastC = 0, dumpC = 0, xoffC[kFRowsPerWarp];
  const float2 *tabC[kFRowsPerWarp][NSLOTS];
  float2 *yC[kFRowsPerWarp];             // FIR: fa.y + (carry + g - N) of this lane at tile 0 of the block; else a.out + g
  int epC[kFRowsPerWarp];
#pragma unroll
  for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
    epA[rr] = epC[rr] = s_row[row0 + rr].ep0; xoffA[rr] = xoffC[rr] = 0; yC[rr] = nullptr;
#pragma unroll
    for (int sl = 0; sl < NSLOTS; ++sl) tabA[rr][sl] = tabC[rr][sl] = a.expj_tables;
  }
  auto block_A = [&](int lb) {
    actA = 0;
#pragma unroll
    for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
      if (lb < lb_run[rr] || lb >= lb_end[rr]) continue;
      actA |= 1u << rr;
      const uint64_t blk = (uint64_t)(s_row[row0 + rr].base + lb);
      while (epA[rr] + 1 < a.nepochs && a.epochs[epA[rr] + 1].first_block <= blk) ++epA[rr];
#pragma unroll
      for (int sl = 0; sl < NSLOTS; ++sl) tabA[rr][sl] = a.expj_tables + (size_t)a.epochs[epA[rr]].table_index[sl] * kNotchN + lane;
      uint32_t lead; block_src(rr, lb, lead);
      xoffA[rr] = (uint32_t)(row0 + rr) * kRawPitch + (lead + lane) * bps;
    }
  };
  auto block_C = [&](int lb) {
    actC = firstC = lastC = dumpC = 0;
#pragma unroll
    for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
      if (lb < lb_run[rr] || lb >= lb_end[rr]) continue;
      const uint64_t blk = (uint64_t)(s_row[row0 + rr].base + lb);
      while (epC[rr] + 1 < a.nepochs && a.epochs[epC[rr] + 1].first_block <= blk) ++epC[rr];   // (also through the warm-up blocks)
      if (lb < lb_own[rr]) continue;                                  // warm-up: the chain ran, nothing is written
      actC |= 1u << rr;
      if (lb == lb_own[rr]) firstC |= 1u << rr;
      if (lb == lb_end[rr] - 1) lastC |= 1u << rr;
#pragma unroll
      for (int sl = 0; sl < NSLOTS; ++sl) tabC[rr][sl] = a.expj_tables + (size_t)a.epochs[epC[rr]].table_index[sl] * kNotchN + lane;
      uint32_t lead; block_src(rr, lb, lead);
      xoffC[rr] = (uint32_t)(row0 + rr) * kRawPitch + (lead + lane) * bps;
      const uint64_t g0 = blk * kNotchN + lane;
      if (FIR) yC[rr] = fa.y + ((int64_t)fa.carry + (int64_t)g0 - N);   // (may point in front of y for the first samples of the first batch: guarded below)
      else yC[rr] = a.out + g0;
      // telemetry: is this block one that cnr_fft / spectrum will look at?  (sorted list, a few entries)
      if (fa.ndump) {
        int lo = 0, hi = fa.ndump;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (fa.dump_blocks[mid] < blk) lo = mid + 1; else hi = mid; }
        if (lo < fa.ndump && fa.dump_blocks[lo] == blk) dumpC |= 1u << rr;
      }
    }
  };
  const bool first_batch = FIR && fa.carry < (uint32_t)N;            // outputs with kk < 0 do not exist

  load(0);
  load(1);
  uint32_t stA = 0, stC = (uint32_t)((kFRawStages - 2) % kFRawStages);   // raw stage of tile s / of tile s - 2
  for (int64_t s = 0; s < total + 2; ++s) {
    load(s + 2);
    cp_async_wait<2>();                          // tile s has landed (this warp loaded its own rows)
    __syncwarp();
    // ------------------------------------------------------------------ A(s)
    if (s < total) {
      const int tib = (int)(s & (kFTilesPerBlock - 1));
      if (tib == 0) block_A((int)(s >> 7));
      if (actA) {
        const unsigned char *stage = s_raw + stA * kStageBytes;
        unsigned char *bkt = s_bk + (uint32_t)(s & 1) * (uint32_t)SM::tile + (uint32_t)row0 * kFPitch + lane * 8;
        const uint32_t eoff = (uint32_t)tib * kFT;
#pragma unroll
        for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
          if (!((actA >> rr) & 1u)) continue;
          const float2 x = row_sample<FMT>(stage + xoffA[rr], 0, a.scale);
#pragma unroll
          for (int sl = 0; sl < NSLOTS; ++sl) {
            const float2 e = __ldg(tabA[rr][sl] + eoff);
            const float br = fmul(fadd(fmul(x.x, e.x), fmul(x.y, e.y)), k);
            const float bi = fmul(fadd(fmul(-x.x, e.y), fmul(x.y, e.x)), k);
            *reinterpret_cast<float2 *>(bkt + (uint32_t)sl * kFRows * kFPitch + rr * kFPitch) = make_float2(br, bi);
          }
        }
      }
    }
    // --------------------------------------------------------------- CD(s - 2)
    const int64_t t = s - 2;
    if (t >= 0) {
      const int tib = (int)(t & (kFTilesPerBlock - 1));
      if (tib == 0) block_C((int)(t >> 7));
      if (actC) {
        const unsigned char *stage = s_raw + stC * kStageBytes;
        const unsigned char *estt = s_est + (uint32_t)(t & 1) * (uint32_t)SM::tile + (uint32_t)row0 * kFPitch + lane * 8;
        const uint32_t eoff = (uint32_t)tib * kFT;
        float2 out[kFRowsPerWarp];
#pragma unroll
        for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
          out[rr] = make_float2(0.f, 0.f);
          if (!((actC >> rr) & 1u)) continue;
          const float2 x = row_sample<FMT>(stage + xoffC[rr], 0, a.scale);
          float outr = x.x, outi = x.y;
#pragma unroll
          for (int sl = 0; sl < NSLOTS; ++sl) {
            const float2 e = __ldg(tabC[rr][sl] + eoff);
            const float2 es = *reinterpret_cast<const float2 *>(estt + (uint32_t)sl * kFRows * kFPitch + rr * kFPitch);
            outr = fsub(outr, fsub(fmul(es.x, e.x), fmul(es.y, e.y)));
            outi = fsub(outi, fadd(fmul(es.x, e.y), fmul(es.y, e.x)));
          }
          if (!unit_gain) { outr = fmul(gain, outr); outi = fmul(gain, outi); }
          out[rr] = make_float2(outr, outi);
        }
        if (dumpC) {                                 // rare: a block the telemetry reads
#pragma unroll
          for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
            if (!((dumpC >> rr) & 1u)) continue;
            const uint64_t blk = (uint64_t)(s_row[row0 + rr].base + (t >> 7));
            int lo = 0, hi = fa.ndump;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (fa.dump_blocks[mid] < blk) lo = mid + 1; else hi = mid; }
            fa.dump[(size_t)lo * kNotchN + tib * kFT + lane] = out[rr];
          }
        }
        if constexpr (!FIR) {
#pragma unroll
          for (int rr = 0; rr < kFRowsPerWarp; ++rr)
            if ((actC >> rr) & 1u) st_stream(yC[rr] + eoff, out[rr]);
        } else {
          // notched tiles of the warp's rows: [previous | current] alternate between the two halves of s_nt
          unsigned char *curb = s_nt + (uint32_t)(t & 1) * (kFRows * kFPitch) + (uint32_t)row0 * kFPitch;
          const int32_t prev_delta = ((t & 1) ? -1 : 1) * (int32_t)(kFRows * kFPitch) + kFT * 8;   // &prev[kFT + j] - &cur[j]
#pragma unroll
          for (int rr = 0; rr < kFRowsPerWarp; ++rr)
            if ((actC >> rr) & 1u) *reinterpret_cast<float2 *>(curb + rr * kFPitch + lane * 8) = out[rr];
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
            if (!((actC >> rr) & 1u)) continue;
            // the output whose newest input is this lane's sample; not for the first N-1 samples of a segment
            // (k_fir_edges), not for outputs in front of the stream
            bool doit = true;
            if (tib == 0) {
              if (((firstC >> rr) & 1u) && lane < L) doit = false;
              if (first_batch && s_row[row0 + rr].base + (t >> 7) == 0 && lane < N) doit = false;
            }
            if (doit) {
              const unsigned char *me = curb + rr * kFPitch + lane * 8;
              float2 acc = make_float2(0.f, 0.f);
              if (fa.real_taps) {
#pragma unroll 5
                for (int i = 0; i < N; ++i) {
                  const float2 v = *reinterpret_cast<const float2 *>(me - i * 8 + ((lane < i) ? prev_delta : 0));
                  const float c = s_taps[i].x;
                  acc.x = fadd(acc.x, fmul(c, v.x)); acc.y = fadd(acc.y, fmul(c, v.y));
                }
              } else {
                for (int i = 0; i < N; ++i) {
                  const float2 v = *reinterpret_cast<const float2 *>(me - i * 8 + ((lane < i) ? prev_delta : 0));
                  const float2 pr = cmul(s_taps[i], v);
                  acc.x = fadd(acc.x, pr.x); acc.y = fadd(acc.y, pr.y);
                }
              }
              st_stream(yC[rr] + eoff, acc);
            }
            // what k_fir_edges needs: the first N-1 and the last N notched samples of the segment
            if (tib == 0 && ((firstC >> rr) & 1u) && lane < L)
              fa.edge[(size_t)s_row[row0 + rr].seg * kNotchEdge + lane] = out[rr];
            if (tib == kFTilesPerBlock - 1 && ((lastC >> rr) & 1u) && lane >= kFT - N)
              fa.edge[(size_t)s_row[row0 + rr].seg * kNotchEdge + kFirFuseMaxTaps + (lane - (kFT - N))] = out[rr];
          }
        }
      }
    }
    stA = (stA + 1 == kFRawStages) ? 0 : stA + 1;
    stC = (stC + 1 == kFRawStages) ? 0 : stC + 1;
    __syncthreads();
  }
}

// The outputs that straddle a segment boundary, and the carry for the next batch.
__global__ void __launch_bounds__(256)
k_fir_edges(NotchFirArgs fa) {
  const NotchApplyArgs &a = fa.n;
  const int N = fa.fir_n, L = N - 1;
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (L > 0) {
    const uint32_t j = gid / (uint32_t)L;
    const int m = (int)(gid % (uint32_t)L);
    if (j < a.nsegs) {
      const uint64_t S = (a.block0 + (uint64_t)j * a.seg_blocks) * (uint64_t)kNotchN;
      const int64_t kk = (int64_t)fa.carry + (int64_t)S + m - N;
      if (kk >= 0) {
        const float2 *first = fa.edge + (size_t)j * kNotchEdge;
        const float2 *last = j ? fa.edge + (size_t)(j - 1) * kNotchEdge + kFirFuseMaxTaps : nullptr;
        float2 acc = make_float2(0.f, 0.f);
        for (int i = 0; i < N; ++i) {
          const int t = m - i;                       // sample S + t
          float2 v;
          if (t >= 0) v = first[t];
          else if (last) v = last[N + t];            // the previous segment ends at S
          else v = fa.carry_in[(int)fa.carry + t];   // batch start: the carried samples
          if (fa.real_taps) {
            const float c = fa.taps[i].x;
            acc.x = fadd(acc.x, fmul(c, v.x)); acc.y = fadd(acc.y, fmul(c, v.y));
          } else {
            const float2 pr = cmul(fa.taps[i], v);
            acc.x = fadd(acc.x, pr.x); acc.y = fadd(acc.y, pr.y);
          }
        }
        fa.y[kk] = acc;
      }
    }
  }
  if (blockIdx.x == 0) {
    // Block 0 holds every thread that read carry_in (j == 0, gid < L <= 31): reads first, then the new carry.
    __syncthreads();
    if ((int)threadIdx.x < N && a.nsegs)
      fa.carry_out[threadIdx.x] = fa.edge[(size_t)(a.nsegs - 1) * kNotchEdge + kFirFuseMaxTaps + threadIdx.x];
  }
}

template <int FMT, int NSLOTS, bool FIR>
cudaError_t launch_nf_t(const NotchFirArgs &fa, const uint32_t *seg_list, uint32_t nlist, const float2 *guess, cudaStream_t st) {
  constexpr size_t smem = FSmem<FMT, NSLOTS, FIR>::total;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static PerDeviceMark configured;
  if (configured.need(1)) {
    cudaError_t e = cudaFuncSetAttribute(k_notch_fir<FMT, NSLOTS, FIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_notch_fir<FMT, NSLOTS, FIR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    configured.commit(1);
  }
  const uint32_t lanes = seg_list ? nlist : fa.n.nsegs;
  if (!lanes) return cudaSuccess;
  k_notch_fir<FMT, NSLOTS, FIR><<<(lanes + kFRows - 1) / kFRows, kFThreads, smem, st>>>(fa, seg_list, nlist, guess);
  return cudaGetLastError();
}

template <int FMT, int NSLOTS>
cudaError_t launch_nf_s(const NotchFirArgs &fa, const uint32_t *seg_list, uint32_t nlist, const float2 *guess, cudaStream_t st) {
  return fa.fir_n > 0 ? launch_nf_t<FMT, NSLOTS, true>(fa, seg_list, nlist, guess, st)
                      : launch_nf_t<FMT, NSLOTS, false>(fa, seg_list, nlist, guess, st);
}

template <int FMT>
cudaError_t launch_nf_f(const NotchFirArgs &fa, const uint32_t *seg_list, uint32_t nlist, const float2 *guess, cudaStream_t st) {
  switch (fa.n.nslots) {
    case 1: return launch_nf_s<FMT, 1>(fa, seg_list, nlist, guess, st);
    case 2: return launch_nf_s<FMT, 2>(fa, seg_list, nlist, guess, st);
    case 3: return launch_nf_s<FMT, 3>(fa, seg_list, nlist, guess, st);
    case 4: return launch_nf_s<FMT, 4>(fa, seg_list, nlist, guess, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t launch_notch_fir(const NotchFirArgs &fa, const uint32_t *seg_list, uint32_t nlist, const float2 *guess, cudaStream_t st) {
  if (fa.n.nsegs == 0 || fa.n.nblocks == 0) return cudaSuccess;
  if (fa.fir_n < 0 || fa.fir_n > kFirFuseMaxTaps) return cudaErrorInvalidValue;
  switch (fa.n.fmt) {
    case 0: return launch_nf_f<0>(fa, seg_list, nlist, guess, st);
    case 1: return launch_nf_f<1>(fa, seg_list, nlist, guess, st);
    case 2: return launch_nf_f<2>(fa, seg_list, nlist, guess, st);
    case 3: return launch_nf_f<3>(fa, seg_list, nlist, guess, st);
    case 4: return launch_nf_f<4>(fa, seg_list, nlist, guess, st);
    default: return launch_nf_f<5>(fa, seg_list, nlist, guess, st);
  }
}

cudaError_t launch_fir_edges(const NotchFirArgs &fa, cudaStream_t st) {
  if (fa.fir_n <= 0 || fa.n.nsegs == 0) return cudaSuccess;
  const uint32_t threads = fa.n.nsegs * (uint32_t)(fa.fir_n > 1 ? fa.fir_n - 1 : 1);
  k_fir_edges<<<(threads + 255) / 256, 256, 0, st>>>(fa);
  return cudaGetLastError();
}

}  // namespace ldvb
