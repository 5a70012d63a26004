// k_notchfir.cu -- auto_notch::process (sdr.h:119-138) and the fir_filter that follows it in leandvb's
// graph (leandvb.cc:296-306 -> :353-384, dsp.h:246-259), in ONE pass: p_notched never exists in HBM.
//
// What is serial in auto_notch is only the one-pole chain  estim = bb*k + estim*(1-k)  (two dependent
// operations per sample and component); the products in front of it (bb*k) and everything behind it
// (out = x - estim*e, the FIR) are data parallel.  k_notch_apply (k_notch.cu) lets every lane do all of it for
// its own segment, so the whole kernel runs at the speed of 1-2 warps per scheduler.  Here the roles are split
// inside a CTA of 16 segments ("rows"):
//
//   worker warps (8 x 2 rows, lane = TWO consecutive samples of a 64-sample tile: 16-byte accesses throughout):
//     LOAD(t+2)  the rows' next raw tile by 16-byte cp.async (each warp loads the rows it works on)
//     A(t)       convert, bk = (x * conj(e)) * k                        -> shared tile bk[t & 1]
//     CD(t-2)    out = gain * (x - sum estim*e); FIR over the row        <- shared tile est[t & 1]
//   chain warp (lane = row):
//     B(t-1)     estim = bk + estim*(1-k), 64 steps per tile, both tiles as conflict-free float4 columns
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

constexpr int kFT = 64;                         // samples per tile: every worker lane handles TWO consecutive samples of a row
constexpr int kFRows = kNotchFirRows;           // 16 segments per CTA (the chain warp uses 16 of its lanes)
constexpr int kFWorkers = 8;                    // worker warps, kFRows / kFWorkers rows each
constexpr int kFRowsPerWarp = kFRows / kFWorkers;
constexpr int kFThreads = (kFWorkers + 1) * 32; // + the chain warp (warp 0)
constexpr int kFRawStages = 5;                  // raw tiles t+2 (in flight) ... t-2 (read again by CD)
constexpr int kFPitch = (kFT + 2) * 8;          // cf32 tiles: 528 B rows, 16 (mod 128): float4 columns are conflict-free
constexpr int kFTilesPerBlock = kNotchN / kFT;
constexpr int kFTileShift = 6;                  // log2(kFTilesPerBlock)
static_assert(kFirFuseMaxTaps <= kFT, "the FIR history of a tile must fit in the previous tile");
static_assert((1 << kFTileShift) == kFTilesPerBlock && kFT == 64, "tile geometry");

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
  if (warp == 0 && lane < kFRows) {
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
    FRow r;
    if (lane < kFRows) r = s_row[lane];
    else { r.base = 0; r.own_begin = r.own_end = r.run_begin = 0; r.seg = 0; r.have = 0; r.ep0 = 0; r.pad = 0; }
    int ep = r.ep0;
    const int lb_run = r.have ? (int)((int64_t)r.run_begin - r.base) : 0;
    const int lb_end = r.have ? (int)((int64_t)r.own_end - r.base) : 0;
    for (int64_t s = 0; s < total + 2; ++s) {
      const int64_t t = s - 1;
      if (t >= 0 && t < total) {
        const int lb = (int)(t >> kFTileShift);
        const int tib = (int)(t & (kFTilesPerBlock - 1));
        const bool active = lb >= lb_run && lb < lb_end;
        if (active) {
          if (tib == 0) {
            // Block start: entry snapshot, epoch switch / resets (sdr.h:97-109).
            const uint64_t blk = (uint64_t)(r.base + lb);
            while (ep + 1 < a.nepochs && a.epochs[ep + 1].first_block <= blk) ++ep;
            if (blk == r.own_begin && a.seg_entry && !forced) {
#pragma unroll
              for (int sl = 0; sl < NSLOTS; ++sl) a.seg_entry[(size_t)r.seg * kNotchMaxSlots + sl] = make_float2(er[sl], ei[sl]);
            }
            if (a.epochs[ep].first_block == blk) {
#pragma unroll
              for (int sl = 0; sl < NSLOTS; ++sl)
                if (a.epochs[ep].reset[sl]) { er[sl] = 0.f; ei[sl] = 0.f; }
            }
          }
          const unsigned char *bk = s_bk + (size_t)(t & 1) * SM::tile + (size_t)lane * kFPitch;
          unsigned char *es = s_est + (size_t)(t & 1) * SM::tile + (size_t)lane * kFPitch;
          // 16 samples (8 float4) per slot at a time; the loads of the next group are issued in front of the stores of
          // this one (the compiler cannot move them there: both go through the same shared array)
          constexpr int G = 8;
#pragma unroll
          for (int sl = 0; sl < NSLOTS; ++sl) {
            const float4 *bp = reinterpret_cast<const float4 *>(bk + (size_t)sl * kFRows * kFPitch);
            float4 *ep4 = reinterpret_cast<float4 *>(es + (size_t)sl * kFRows * kFPitch);
            float4 b[G], nb[G];
#pragma unroll
            for (int q = 0; q < G; ++q) b[q] = bp[q];
#pragma unroll
            for (int c = 0; c < kFT / 2 / G; ++c) {
              if (c + 1 < kFT / 2 / G) {
#pragma unroll
                for (int q = 0; q < G; ++q) nb[q] = bp[(c + 1) * G + q];
              }
#pragma unroll
              for (int q = 0; q < G; ++q) {
                const float r0 = fadd(b[q].x, fmul(er[sl], omk)), i0 = fadd(b[q].y, fmul(ei[sl], omk));
                er[sl] = fadd(b[q].z, fmul(r0, omk)); ei[sl] = fadd(b[q].w, fmul(i0, omk));
                ep4[c * G + q] = make_float4(r0, i0, er[sl], ei[sl]);
              }
#pragma unroll
              for (int q = 0; q < G; ++q) b[q] = nb[q];
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
  // tables are in force, where do its outputs go -- is worked out once per 4096-sample block (64 tiles) and kept
  // in registers; a tile step then costs a handful of instructions per row besides the arithmetic.  A lane handles
  // the samples 2*lane and 2*lane + 1 of a row-tile: 16-byte shared-memory and table accesses.
  const int w = warp - 1;
  const int row0 = w * kFRowsPerWarp;
  const int N = FIR ? fa.fir_n : 0, L = N > 0 ? N - 1 : 0;
  constexpr int kPieces = kRawPitch / 16;                             // 16-byte pieces of a staged row, incl. the extra one
  constexpr int kLoadIters = (kFRowsPerWarp * kPieces + 31) / 32;
  constexpr uint32_t kStageBytes = (uint32_t)kFRows * kRawPitch;

  int lb_run[kFRowsPerWarp], lb_own[kFRowsPerWarp], lb_end[kFRowsPerWarp];
#pragma unroll
  for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
    const FRow &r = s_row[row0 + rr];
    lb_run[rr] = r.have ? (int)((int64_t)r.run_begin - r.base) : 0;
    lb_own[rr] = r.have ? (int)((int64_t)r.own_begin - r.base) : 0;
    lb_end[rr] = r.have ? (int)((int64_t)r.own_end - r.base) : 0;
  }
  auto block_src = [&](int rr, int lb, uint32_t &lead) -> const unsigned char * {
    uint64_t idx = (uint64_t)(s_row[row0 + rr].base + lb) * kNotchN;
    const unsigned char *part = static_cast<const unsigned char *>(a.src.head);
    lead = 0;
    if (a.src.main && idx >= a.src.c0) { part = static_cast<const unsigned char *>(a.src.main); idx -= a.src.c0; lead = lead_main; }
    return part + (idx - lead) * bps;
  };

  // ---- LOAD state: this lane's pieces of the warp's rows
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
        const FRow &r = s_row[row0 + rr];          // (rr is a per-lane value here)
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
      if (tib == 0) load_block((int)(t >> kFTileShift));
      unsigned char *stage = s_raw + stL * kStageBytes;
      const uint32_t toff = (uint32_t)tib * kFT * bps;
#pragma unroll
      for (int i = 0; i < kLoadIters; ++i)
        if ((ld_ok >> i) & 1u) cp_async16(stage + ld_dst[i], ld_src[i] + toff);
    }
    stL = (stL + 1 == kFRawStages) ? 0 : stL + 1;
    cp_async_commit();
  };

  // ---- A / CD state
  uint32_t actA = 0, xoffA[kFRowsPerWarp];
  const float2 *tabA[kFRowsPerWarp][NSLOTS];
  int epA[kFRowsPerWarp];
  uint32_t actC = 0, firstC = 0, lastC = 0, dumpC = 0, xoffC[kFRowsPerWarp];
  const float2 *tabC[kFRowsPerWarp][NSLOTS];
  float2 *yC[kFRowsPerWarp];             // FIR: fa.y + (carry + g - N) of this lane's first sample at tile 0 of the block; else a.out + g
  float2 *dmpC[kFRowsPerWarp];
  int epC[kFRowsPerWarp];
#pragma unroll
  for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
    epA[rr] = epC[rr] = s_row[row0 + rr].ep0; xoffA[rr] = xoffC[rr] = 0; yC[rr] = nullptr; dmpC[rr] = nullptr;
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
      for (int sl = 0; sl < NSLOTS; ++sl) tabA[rr][sl] = a.expj_tables + (size_t)a.epochs[epA[rr]].table_index[sl] * kNotchN + 2 * lane;
      uint32_t lead; block_src(rr, lb, lead);
      xoffA[rr] = (uint32_t)(row0 + rr) * kRawPitch + (lead + 2 * lane) * bps;
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
      for (int sl = 0; sl < NSLOTS; ++sl) tabC[rr][sl] = a.expj_tables + (size_t)a.epochs[epC[rr]].table_index[sl] * kNotchN + 2 * lane;
      uint32_t lead; block_src(rr, lb, lead);
      xoffC[rr] = (uint32_t)(row0 + rr) * kRawPitch + (lead + 2 * lane) * bps;
      const uint64_t g0 = blk * kNotchN + 2 * lane;
      if (FIR) yC[rr] = fa.y + ((int64_t)fa.carry + (int64_t)g0 - N);   // (may point in front of y for the first samples of the first batch: guarded below)
      else yC[rr] = a.out + g0;
      // telemetry: is this block one that cnr_fft / spectrum will look at?  (sorted list, a few entries)
      if (fa.ndump) {
        int lo = 0, hi = fa.ndump;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (fa.dump_blocks[mid] < blk) lo = mid + 1; else hi = mid; }
        if (lo < fa.ndump && fa.dump_blocks[lo] == blk) { dumpC |= 1u << rr; dmpC[rr] = fa.dump + (size_t)lo * kNotchN + 2 * lane; }
      }
    }
  };
  const bool first_batch = FIR && fa.carry < (uint32_t)N;            // outputs with kk < 0 do not exist
  // two consecutive samples of a staged raw row
  auto raw_pair = [&](const unsigned char *p, float2 &x0, float2 &x1) {
    if (FMT >= 4 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      const float4 v = *reinterpret_cast<const float4 *>(p);
      x0 = make_float2(v.x, v.y); x1 = make_float2(v.z, v.w);
      if (FMT == 4 && a.scale != 1.0f) { x0.x = fmul(x0.x, a.scale); x0.y = fmul(x0.y, a.scale); x1.x = fmul(x1.x, a.scale); x1.y = fmul(x1.y, a.scale); }
    } else {
      x0 = row_sample<FMT>(p, 0, a.scale); x1 = row_sample<FMT>(p, 1, a.scale);
    }
  };
  auto store_pair = [&](float2 *dst, float2 v0, float2 v1) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) st_stream(reinterpret_cast<float4 *>(dst), make_float4(v0.x, v0.y, v1.x, v1.y));
    else { st_stream(dst, v0); st_stream(dst + 1, v1); }
  };

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
      if (tib == 0) block_A((int)(s >> kFTileShift));
      if (actA) {
        const unsigned char *stage = s_raw + stA * kStageBytes;
        unsigned char *bkt = s_bk + (uint32_t)(s & 1) * (uint32_t)SM::tile + (uint32_t)row0 * kFPitch + lane * 16;
        const uint32_t eoff = (uint32_t)tib * kFT;
#pragma unroll
        for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
          if (!((actA >> rr) & 1u)) continue;
          float2 x0, x1;
          raw_pair(stage + xoffA[rr], x0, x1);
#pragma unroll
          for (int sl = 0; sl < NSLOTS; ++sl) {
            const float4 e = __ldg(reinterpret_cast<const float4 *>(tabA[rr][sl] + eoff));
            float4 b;
            b.x = fmul(fadd(fmul(x0.x, e.x), fmul(x0.y, e.y)), k);
            b.y = fmul(fadd(fmul(-x0.x, e.y), fmul(x0.y, e.x)), k);
            b.z = fmul(fadd(fmul(x1.x, e.z), fmul(x1.y, e.w)), k);
            b.w = fmul(fadd(fmul(-x1.x, e.w), fmul(x1.y, e.z)), k);
            *reinterpret_cast<float4 *>(bkt + (uint32_t)sl * kFRows * kFPitch + rr * kFPitch) = b;
          }
        }
      }
    }
    // --------------------------------------------------------------- CD(s - 2)
    const int64_t t = s - 2;
    if (t >= 0) {
      const int tib = (int)(t & (kFTilesPerBlock - 1));
      if (tib == 0) block_C((int)(t >> kFTileShift));
      if (actC) {
        const unsigned char *stage = s_raw + stC * kStageBytes;
        const unsigned char *estt = s_est + (uint32_t)(t & 1) * (uint32_t)SM::tile + (uint32_t)row0 * kFPitch + lane * 16;
        const uint32_t eoff = (uint32_t)tib * kFT;
        float2 o0[kFRowsPerWarp], o1[kFRowsPerWarp];
#pragma unroll
        for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
          o0[rr] = o1[rr] = make_float2(0.f, 0.f);
          if (!((actC >> rr) & 1u)) continue;
          float2 x0, x1;
          raw_pair(stage + xoffC[rr], x0, x1);
          float r0 = x0.x, i0 = x0.y, r1 = x1.x, i1 = x1.y;
#pragma unroll
          for (int sl = 0; sl < NSLOTS; ++sl) {
            const float4 e = __ldg(reinterpret_cast<const float4 *>(tabC[rr][sl] + eoff));
            const float4 es = *reinterpret_cast<const float4 *>(estt + (uint32_t)sl * kFRows * kFPitch + rr * kFPitch);
            r0 = fsub(r0, fsub(fmul(es.x, e.x), fmul(es.y, e.y)));
            i0 = fsub(i0, fadd(fmul(es.x, e.y), fmul(es.y, e.x)));
            r1 = fsub(r1, fsub(fmul(es.z, e.z), fmul(es.w, e.w)));
            i1 = fsub(i1, fadd(fmul(es.z, e.w), fmul(es.w, e.z)));
          }
          if (!unit_gain) { r0 = fmul(gain, r0); i0 = fmul(gain, i0); r1 = fmul(gain, r1); i1 = fmul(gain, i1); }
          o0[rr] = make_float2(r0, i0); o1[rr] = make_float2(r1, i1);
          if ((dumpC >> rr) & 1u)                  // rare: a block the telemetry reads
            *reinterpret_cast<float4 *>(dmpC[rr] + eoff) = make_float4(r0, i0, r1, i1);
        }
        if constexpr (!FIR) {
#pragma unroll
          for (int rr = 0; rr < kFRowsPerWarp; ++rr)
            if ((actC >> rr) & 1u) store_pair(yC[rr] + eoff, o0[rr], o1[rr]);
        } else {
          // notched tiles of the warp's rows: [previous | current] alternate between the two halves of s_nt
          unsigned char *curb = s_nt + (uint32_t)(t & 1) * (kFRows * kFPitch) + (uint32_t)row0 * kFPitch + lane * 16;
          const int32_t prev_delta = ((t & 1) ? -1 : 1) * (int32_t)(kFRows * kFPitch) + kFT * 8;   // &prev[kFT + j] - &cur[j]
#pragma unroll
          for (int rr = 0; rr < kFRowsPerWarp; ++rr)
            if ((actC >> rr) & 1u) *reinterpret_cast<float4 *>(curb + rr * kFPitch) = make_float4(o0[rr].x, o0[rr].y, o1[rr].x, o1[rr].y);
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < kFRowsPerWarp; ++rr) {
            if (!((actC >> rr) & 1u)) continue;
            // Outputs whose newest inputs are this lane's two samples (2 lane, 2 lane + 1): a sliding pair over the taps,
            //   y1 = sum_i c_i v[2l+1-i],  y0 = sum_i c_i v[2l-i]   (accumulated in the reference's order, i = 0 .. N-1)
            const unsigned char *me = curb + rr * kFPitch;      // &cur[2 lane]
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
            float2 prev = o1[rr], cur = o0[rr];
            if (fa.real_taps) {
#pragma unroll 5
              for (int i = 0; i < N; ++i) {
                const float c = s_taps[i].x;
                a1.x = fadd(a1.x, fmul(c, prev.x)); a1.y = fadd(a1.y, fmul(c, prev.y));
                a0.x = fadd(a0.x, fmul(c, cur.x));  a0.y = fadd(a0.y, fmul(c, cur.y));
                prev = cur;
                const int j = 2 * lane - (i + 1);                  // the next older sample
                cur = *reinterpret_cast<const float2 *>(me - (i + 1) * 8 + ((j < 0) ? prev_delta : 0));
              }
            } else {
              for (int i = 0; i < N; ++i) {
                const float2 c = s_taps[i];
                const float2 p1 = cmul(c, prev), p0 = cmul(c, cur);
                a1.x = fadd(a1.x, p1.x); a1.y = fadd(a1.y, p1.y);
                a0.x = fadd(a0.x, p0.x); a0.y = fadd(a0.y, p0.y);
                prev = cur;
                const int j = 2 * lane - (i + 1);
                cur = *reinterpret_cast<const float2 *>(me - (i + 1) * 8 + ((j < 0) ? prev_delta : 0));
              }
            }
            // not for the first N-1 samples of a segment (k_fir_edges), not for outputs in front of the stream
            bool ok0 = true, ok1 = true;
            if (tib == 0) {
              if ((firstC >> rr) & 1u) { ok0 = 2 * lane >= L; ok1 = 2 * lane + 1 >= L; }
              if (first_batch && s_row[row0 + rr].base + (t >> kFTileShift) == 0) { ok0 = ok0 && 2 * lane >= N; ok1 = ok1 && 2 * lane + 1 >= N; }
            }
            float2 *yp = yC[rr] + eoff;
            if (ok0 && ok1) store_pair(yp, a0, a1);
            else { if (ok0) st_stream(yp, a0); if (ok1) st_stream(yp + 1, a1); }
            // what k_fir_edges needs: the first N-1 and the last N notched samples of the segment
            if (tib == 0 && ((firstC >> rr) & 1u)) {
              float2 *ed = fa.edge + (size_t)s_row[row0 + rr].seg * kNotchEdge;
              if (2 * lane < L) ed[2 * lane] = o0[rr];
              if (2 * lane + 1 < L) ed[2 * lane + 1] = o1[rr];
            }
            if (tib == kFTilesPerBlock - 1 && ((lastC >> rr) & 1u)) {
              float2 *ed = fa.edge + (size_t)s_row[row0 + rr].seg * kNotchEdge + kFirFuseMaxTaps - (kFT - N);
              if (2 * lane >= kFT - N) ed[2 * lane] = o0[rr];
              if (2 * lane + 1 >= kFT - N) ed[2 * lane + 1] = o1[rr];
            }
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
