// k_notch.cu -- K2: auto_notch<f32> (sdr.h:46-154).
//
// detect(): every 1024*4096 input samples the reference runs a 4096-point
//   radix-2 inverse FFT (cfft_engine::inplace, dsp.h:78-110) on the current
//   block, takes hypotf() of every bin and picks the nslots strongest peaks
//   (first maximum wins, neighbours zeroed; sdr.h:76-118).  One CTA per detect
//   point reproduces the butterfly order and rounding exactly: butterflies of a
//   stage are independent, so they run in parallel with the same per-element
//   arithmetic; hypotf is evaluated like glibc does, in double.
// process(): per sample and slot a one-pole estimate of the birdie
//   estim = bb*k + estim*(1-k), bb = x*conj(e[n]), out = x - estim*e[n]
//   (sdr.h:119-138).  The recurrence is serial in time and float addition is not
//   associative, so an exact parallel prefix does not exist.  The stream is cut
//   into segments that run concurrently; a segment that does not start at a
//   known-exact state (batch start or a reset at a detect point) starts
//   `warm_blocks` earlier from a zero estimate: the influence of the start value
//   decays as 0.998^n and vanishes below one ulp, after which both trajectories
//   round identically for ever.  Every segment records its state at entry and
//   exit; the host checks entry(j+1) == exit(j) bit for bit and re-runs the rare
//   segment whose warm-up had not merged (exact by induction over segments).
#include "common.cuh"
#include <cstdlib>

#include "kernels.h"
#include "notch_common.cuh"

namespace ldvb {

namespace {

__device__ __forceinline__ float2 load_sample(const RawSrc &src, int fmt, uint64_t idx, float scale) {
  const void *raw = src.head;
  if (src.main && idx >= src.c0) { raw = src.main; idx -= src.c0; }
  switch (fmt) {
    case 0: {
      uchar2 v = reinterpret_cast<const uchar2 *>(raw)[idx];
      return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128));
    }
    case 1: {
      char2 v = reinterpret_cast<const char2 *>(raw)[idx];
      return make_float2((float)(int)v.x, (float)(int)v.y);
    }
    case 2: {
      ushort2 v = reinterpret_cast<const ushort2 *>(raw)[idx];
      return make_float2((float)((int)v.x - 32768), (float)((int)v.y - 32768));
    }
    case 3: {
      short2 v = reinterpret_cast<const short2 *>(raw)[idx];
      return make_float2((float)(int)v.x, (float)(int)v.y);
    }
    case 4: {
      float2 v = __ldg(reinterpret_cast<const float2 *>(raw) + idx);
      return make_float2(fmul(v.x, scale), fmul(v.y, scale));
    }
    default:
      return __ldg(reinterpret_cast<const float2 *>(raw) + idx);
  }
}

// --------------------------------------------------------------------- detect
__global__ void __launch_bounds__(1024)
k_notch_detect(NotchDetectArgs a) {
  __shared__ float2 d[kNotchN];
  float *amp = reinterpret_cast<float *>(d);  // reused once the spectrum is in registers
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  const int tid = threadIdx.x;
  const uint64_t base = a.block_index[blockIdx.x] * (uint64_t)kNotchN;
  // Load with the bit-reversal permutation applied (dsp.h:79-83 swaps i <-> rev(i)).
  for (int i = tid; i < kNotchN; i += 1024) {
    const int r = (int)(__brev((unsigned)i) >> 20);
    d[r] = load_sample(a.src, a.fmt, base + i, a.scale);
  }
  __syncthreads();
  // Danielson-Lanczos stages (dsp.h:85-102), twiddle om[k*dom] from the host table.
  for (int s = 0; s < 12; ++s) {
    const int hbs = 1 << s, dom = 1 << (11 - s);
    for (int b = tid; b < kNotchN / 2; b += 1024) {
      const int j = b >> s, k = b & (hbs - 1);
      const int pidx = j * hbs * 2 + k, qidx = pidx + hbs;
      const float2 w = a.twiddle_rev[k * dom];
      const float2 q = d[qidx], p = d[pidx];
      const float xr = fsub(fmul(w.x, q.x), fmul(w.y, q.y));
      const float xi = fadd(fmul(w.x, q.y), fmul(w.y, q.x));
      d[qidx] = make_float2(fsub(p.x, xr), fsub(p.y, xi));
      d[pidx] = make_float2(fadd(p.x, xr), fadd(p.y, xi));
    }
    __syncthreads();
  }
  const float invn = 1.0f / kNotchN;  // dsp.h:104-109
  float my_amp[kNotchN / 1024];
  for (int q = 0; q < kNotchN / 1024; ++q) {
    const int i = tid + 1024 * q;
    const float re = fmul(d[i].x, invn), im = fmul(d[i].y, invn);
    // glibc hypotf: (float)sqrt((double)x*x + (double)y*y)
    const double s2 = __dadd_rn(__dmul_rn((double)re, (double)re), __dmul_rn((double)im, (double)im));
    my_amp[q] = __double2float_rn(__dsqrt_rn(s2));
  }
  __syncthreads();
  for (int q = 0; q < kNotchN / 1024; ++q) amp[tid + 1024 * q] = my_amp[q];
  __syncthreads();
  for (int slot = 0; slot < a.nslots; ++slot) {
    // argmax, first maximum wins (strict '>' in the reference's forward scan).
    float bv = -1.f; int bi = 0;
    for (int i = tid; i < kNotchN; i += 1024) {
      const float v = amp[i];
      if (v > bv) { bv = v; bi = i; }
    }
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
      bv = s_val[tid]; bi = s_idx[tid];
      for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (tid == 0) {
        a.bins_out[blockIdx.x * a.nslots + slot] = bi;
        amp[bi] = 0;
        if (bi - 1 >= 0) amp[bi - 1] = 0;
        if (bi + 1 < kNotchN) amp[bi + 1] = 0;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------- apply
// Start-state guess.  The estimate forgets its past with a time constant of
// 1/k = 500 samples, so the state at any point is, to ~1e-7 relative, the
// exponentially weighted sum of the last few thousand inputs:
//   estim(P) ~ sum_{m>=1} (bb[P-m]*k) * (1-k)^(m-1)
// One warp evaluates that sum for one segment (parallel over m).  The guess is
// NOT exact; it only has to land within a few ulps so that the exact serial
// run that follows merges with the true trajectory inside the one warm-up block
// (measured: median 1000 samples, max < 3000).  Every segment is
// still verified against its predecessor and re-run when it did not merge.

// Block sums: S_b = sum over the samples i of block b of bb[i]*k*(1-k)^(4095-i), i.e. what
// block b alone contributes to the estimate at its end.  One CTA per block, every sample of
// the stream is read once (coalesced); the start state of a segment is then assembled from
// the two block sums in front of its warm-up (guess_from_sums).
template <int FMT, int NSLOTS>
__global__ void __launch_bounds__(128)
k_notch_guess(NotchApplyArgs a, uint64_t first_block, float2 *sums /* [nblocks][kNotchMaxSlots] */, const float *weights) {
  __shared__ float2 part[4][NSLOTS];
  const uint64_t b = first_block + blockIdx.x;
  if (b >= a.nblocks) return;
  if (a.seg_blocks > 2) {
    // Only the two blocks in front of a segment's warm-up are ever asked for (guess_from_sums): with long
    // segments most blocks are skipped (b + warm + 2 >= block0 by the choice of first_block).
    if ((b + a.warm_blocks + 2 - a.block0) % a.seg_blocks > 1) return;
  }
  int ep = 0;
  while (ep + 1 < a.nepochs && a.epochs[ep + 1].first_block <= b) ++ep;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float accr[NSLOTS], acci[NSLOTS];
  const float2 *tab[NSLOTS];
#pragma unroll
  for (int s = 0; s < NSLOTS; ++s) { accr[s] = 0.f; acci[s] = 0.f; tab[s] = a.expj_tables + (size_t)a.epochs[ep].table_index[s] * kNotchN; }
  // The block lies in ONE part of the two-part stream unless it is the block that holds the carry boundary.
  const uint64_t base = b * (uint64_t)kNotchN;
  const bool split = a.src.main && base < a.src.c0 && base + kNotchN > a.src.c0;
  const float k = a.k;
  if (FMT >= 4 && !split) {
    // cf32: two samples per 16-byte load when the block starts on a 16-byte boundary of its part, all loads of an
    // iteration group issued before their first use (plain streaming, bound by the one pass over the samples).
    const bool in_main = a.src.main && base >= a.src.c0;
    const float2 *xs = in_main ? reinterpret_cast<const float2 *>(a.src.main) + (base - a.src.c0)
                               : reinterpret_cast<const float2 *>(a.src.head) + base;
    const bool al16 = (reinterpret_cast<uintptr_t>(xs) & 15u) == 0;
#pragma unroll 4
    for (int j = 0; j < kNotchN / 256; ++j) {
      const int i = 2 * (j * 128 + tid);
      float2 x0, x1;
      if (al16) { const float4 v = __ldcs(reinterpret_cast<const float4 *>(xs + i)); x0 = make_float2(v.x, v.y); x1 = make_float2(v.z, v.w); }
      else { x0 = __ldcs(xs + i); x1 = __ldcs(xs + i + 1); }
      if (FMT == 4 && a.scale != 1.0f) { x0.x = fmul(x0.x, a.scale); x0.y = fmul(x0.y, a.scale); x1.x = fmul(x1.x, a.scale); x1.y = fmul(x1.y, a.scale); }
      // weights[m] = (1-k)^m, m = 4095 - i: the pair (i, i+1) reads weights[4094 - i], weights[4095 - i]
      const float2 wv = __ldg(reinterpret_cast<const float2 *>(weights + (kNotchN - 2 - i)));
      const float w0 = wv.y * k, w1 = wv.x * k;
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) {
        const float4 e = __ldg(reinterpret_cast<const float4 *>(tab[s] + i));
        accr[s] += (x0.x * e.x + x0.y * e.y) * w0 + (x1.x * e.z + x1.y * e.w) * w1;
        acci[s] += (-x0.x * e.y + x0.y * e.x) * w0 + (-x1.x * e.w + x1.y * e.z) * w1;
      }
    }
  } else {
#pragma unroll 4
    for (int j = 0; j < kNotchN / 128; ++j) {
      const int i = j * 128 + tid;
      const float2 x = ld_raw<FMT>(a.src, base + i, a.scale);
      const float w = __ldg(weights + (kNotchN - 1 - i)) * k;
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) {
        const float2 e = __ldg(tab[s] + i);
        accr[s] += (x.x * e.x + x.y * e.y) * w;
        acci[s] += (-x.x * e.y + x.y * e.x) * w;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < NSLOTS; ++s) {
    for (int o = 16; o; o >>= 1) {
      accr[s] += __shfl_xor_sync(0xffffffffu, accr[s], o);
      acci[s] += __shfl_xor_sync(0xffffffffu, acci[s], o);
    }
    if (lane == 0) part[warp][s] = make_float2(accr[s], acci[s]);
  }
  __syncthreads();
  if (tid < NSLOTS) {
    float2 t = part[0][tid];
    for (int w = 1; w < 4; ++w) { t.x += part[w][tid].x; t.y += part[w][tid].y; }
    sums[b * kNotchMaxSlots + tid] = t;
  }
}

// One lane = one segment.  The 32 lanes of a warp walk 32 segments in lock step;
// per tile the warp stages every lane's next 64 raw samples in a private
// shared-memory row (cooperative 16-byte cp.async copies, contiguous within a
// row, double buffered); lanes convert on the fly and stream their results out.
#ifndef LDVB_NOTCH_TILE
#define LDVB_NOTCH_TILE 32
#endif
#ifndef LDVB_NOTCH_STAGES
#define LDVB_NOTCH_STAGES 2
#endif
#ifndef LDVB_NOTCH_WARPS
#define LDVB_NOTCH_WARPS 2
#endif
constexpr int kNTile = LDVB_NOTCH_TILE;            // 16, 32 or 64 samples per staged tile
constexpr int kNPitch = (kNTile + 2) * 8;          // row pitch: (tile + 2) cf32, = 16 (mod 128) bytes
constexpr int kNStages = LDVB_NOTCH_STAGES;        // tiles in flight per lane: kNStages - 1 ahead of the one in use
constexpr int kNWarps = LDVB_NOTCH_WARPS;
static_assert(kNPitch % 128 == 16 && kNotchN % kNTile == 0 && kNStages >= 2, "row geometry");
// dynamic shared memory: [input rows: warps x stages x 32 x pitch | table tiles: warps x stages x slots x tile | output tiles]
// (two table sets per stage and slot, see k_notch_apply; sized by the number of slots so that the one-slot kernel
//  still fits four CTAs per SM)
template <int NSLOTS> struct NotchSmem {
  static constexpr size_t in = (size_t)kNWarps * kNStages * 32 * kNPitch + (size_t)kNWarps * kNStages * 2 * NSLOTS * kNTile * 8;
  static constexpr size_t total = in + (size_t)kNWarps * 32 * kNPitch;   // + one output tile per warp
};

template <int FMT, int NSLOTS>
__global__ void __launch_bounds__(kNWarps * 32)
k_notch_apply(NotchApplyArgs a, const uint32_t *seg_list, uint32_t nlist, const float2 *guess) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t warp_global = blockIdx.x * kNWarps + warp;
  // Normal mode: lane g owns segment g.  Repair mode (seg_list): lane g re-runs
  // segment seg_list[g] exactly, from the exit state of its predecessor.
  const bool repair = (seg_list != nullptr);
  uint32_t seg; bool have;
  const uint32_t g = warp_global * 32 + lane;
  if (repair) { have = g < nlist; seg = have ? seg_list[g] : 0; }
  else { seg = g; have = seg < a.nsegs; }
  const float2 *forced_entry = (repair && have) ? a.seg_exit + (size_t)(seg - 1) * kNotchMaxSlots : nullptr;

  SegPlan p;
  p.own_begin = p.own_end = p.run_begin = 0; p.epoch = 0; p.start_kind = 0;
  float er[NSLOTS], ei[NSLOTS];
#pragma unroll
  for (int s = 0; s < NSLOTS; ++s) { er[s] = 0.f; ei[s] = 0.f; }
  int ep = 0;
  if (have) {
    p = plan_segment(a, seg);
    ep = p.epoch;
    if (forced_entry) {
      p.run_begin = p.own_begin;
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) { er[s] = forced_entry[s].x; ei[s] = forced_entry[s].y; }
    } else if (p.start_kind == 0) {
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) { er[s] = a.state_in->slot[s].est_re; ei[s] = a.state_in->slot[s].est_im; }
    } else if (p.start_kind == 2) {
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) { const float2 g = guess_from_sums(a, p, guess, s); er[s] = g.x; ei[s] = g.y; }
    }
  }
  const uint64_t own_begin = p.own_begin, own_end = p.own_end, run_begin = p.run_begin;
  // Common iteration space: local block i -> block = base + i.
  int64_t base; uint64_t iters;
  if (repair) { base = (int64_t)run_begin; iters = a.seg_blocks; }
  else { base = (int64_t)(a.block0 + (uint64_t)seg * a.seg_blocks) - (int64_t)a.warm_blocks; iters = (uint64_t)a.warm_blocks + a.seg_blocks; }

  // This lane's private row in each stage.
  const uint32_t row_off = (uint32_t)(warp * kNStages * 32 * kNPitch + lane * kNPitch);

  constexpr uint32_t bps = (FMT <= 1) ? 2u : (FMT <= 3 ? 4u : 8u);
  constexpr uint32_t align_elems = 16 / bps;
  constexpr int kTilesPerBlock = kNotchN / kNTile;
  const uint64_t total_tiles = iters * kTilesPerBlock;
  // Where sample `idx` of the two-part stream lives, aligned down to 16 bytes.
  auto locate = [&](uint64_t idx, const unsigned char *&src, uint32_t &lead) {
    const unsigned char *part = static_cast<const unsigned char *>(a.src.head);
    if (a.src.main && idx >= a.src.c0) { part = static_cast<const unsigned char *>(a.src.main); idx -= a.src.c0; }
    const uint64_t al = idx & ~(uint64_t)(align_elems - 1);
    lead = (uint32_t)(idx - al);
    src = part + al * bps;
  };
  constexpr int kRowChunks = (int)(kNTile * bps / 16u);      // whole 16-byte pieces of a row
  constexpr int kRowsPerCopy = 32 / kRowChunks;              // rows covered by one warp-wide copy
  static_assert(kRowChunks >= 1 && 32 % kRowChunks == 0, "row geometry");
  // The e^{j theta} tile of the warp: the 32 lanes walk 32 different blocks but sit at the SAME
  // offset inside them, so they all need the same kNTile table entries per slot.  When the
  // active lanes agree on the table (always, except in a warp that straddles a bin change) the
  // tile is staged once per warp with the rows (kNTile * 8 bytes per slot, 16 B per lane) and
  // read back as shared-memory broadcasts: the table leaves the dependent-load path (ncu: the
  // products waiting on these LDGs were 45 % of the kernel's stall samples).
  // Up to TWO table sets per stage: a warp that straddles a bin change (an epoch boundary, 30 per 128 M samples on
  // a flat spectrum) has lanes on the old tables and lanes on the new ones; with one set it fell back to per-lane
  // global loads for all its three blocks and, the kernel being one wave of latency-bound warps, set the kernel's
  // duration (1.39 ms against 1.00 ms, continuous-stream bench of round 2).
  float2 *wtab = reinterpret_cast<float2 *>(smem + (size_t)kNWarps * kNStages * 32 * kNPitch) +
                 (size_t)warp * kNStages * 2 * NSLOTS * kNTile;
  uint32_t set_mask = 0;              // bit st: this LANE reads set 1 of stage st
  uint32_t staged_mask = 0;           // bit st: stage st holds a valid table tile (warp uniform)
  int epi = p.epoch;                  // epoch cursor of the issue stream
  uint32_t tixi[NSLOTS];
#pragma unroll
  for (int s = 0; s < NSLOTS; ++s) tixi[s] = have ? a.epochs[epi].table_index[s] : 0u;
  // Each lane fetches ITS row with 16-byte asynchronous copies (LDGSTS).
  auto issue = [&](uint64_t tile) {
    const int st = (int)(tile % kNStages);
    const int64_t blk = base + (int64_t)(tile / kTilesPerBlock);
    const int tibi = (int)(tile % kTilesPerBlock);
    const bool active = have && blk >= (int64_t)run_begin && blk < (int64_t)own_end;
    const unsigned char *src = nullptr; uint32_t lead = 0;
    if (active) {
      locate((uint64_t)blk * kNotchN + (uint64_t)tibi * kNTile, src, lead);
      if (tibi == 0) {
        const int before = epi;
        while (epi + 1 < a.nepochs && a.epochs[epi + 1].first_block <= (uint64_t)blk) ++epi;
        if (epi != before) {
#pragma unroll
          for (int s = 0; s < NSLOTS; ++s) tixi[s] = a.epochs[epi].table_index[s];
        }
      }
    }
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (act) {
      // Rows are fetched by the WARP, not by their owners: kRowChunks lanes cover one row with
      // consecutive 16-byte pieces, so one LDGSTS touches 32 / kRowChunks rows (2 x 128-byte lines
      // each) instead of 32 different lines -- the per-lane version kept the L1 -> crossbar request
      // path 57 % busy (ncu, round 1), which is what bounded the kernel.  Row addresses travel by
      // shuffle; the odd 16-byte piece of a row that does not start on a 16-byte boundary is
      // fetched by its owner.
      const uint32_t slo = (uint32_t)reinterpret_cast<uintptr_t>(src), shi = (uint32_t)(reinterpret_cast<uintptr_t>(src) >> 32);
      const int q = lane % kRowChunks, sub = lane / kRowChunks;
      unsigned char *stage_base = smem + (size_t)warp * kNStages * 32 * kNPitch + (size_t)st * 32 * kNPitch;
#pragma unroll
      for (int r = 0; r < 32; r += kRowsPerCopy) {
        const int row = r + sub;
        const uint32_t lo = __shfl_sync(0xffffffffu, slo, row), hi = __shfl_sync(0xffffffffu, shi, row);
        if ((act >> row) & 1u) {
          const unsigned char *rs = reinterpret_cast<const unsigned char *>(((uintptr_t)hi << 32) | lo);
          cp_async16(stage_base + (size_t)row * kNPitch + q * 16, rs + q * 16);
        }
      }
      if (active && (lead + kNTile) * bps > (uint32_t)kRowChunks * 16u)
        cp_async16(stage_base + (size_t)lane * kNPitch + kRowChunks * 16, src + kRowChunks * 16);
    }
    // Which tables do the active lanes use?  One set (always, except in a warp that straddles a bin change) or two
    // are staged; three or more (never seen) fall back to per-lane global loads.  (All lanes of the warp are here.)
    bool stage_ok = act != 0;
    bool two = false;
    uint32_t lead_tix[NSLOTS], alt_tix[NSLOTS];
    bool mine_alt = false;
    if (act) {
      const int leader = __ffs(act) - 1;
      bool same = true;
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) {
        lead_tix[s] = __shfl_sync(0xffffffffu, tixi[s], leader);
        alt_tix[s] = lead_tix[s];
        same = same && (!active || tixi[s] == lead_tix[s]);
      }
      const unsigned diff = __ballot_sync(0xffffffffu, !same);
      if (diff) {
        const int l2 = __ffs(diff) - 1;
        bool same2 = true;
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) {
          alt_tix[s] = __shfl_sync(0xffffffffu, tixi[s], l2);
          same2 = same2 && (tixi[s] == alt_tix[s]);
        }
        mine_alt = !same;
        two = true;
        stage_ok = __all_sync(0xffffffffu, same || same2);
      }
    }
    if (stage_ok) {
      staged_mask |= 1u << st;
      if (mine_alt) set_mask |= 1u << st; else set_mask &= ~(1u << st);
      constexpr int kChunks = kNTile * 8 / 16;       // 16-byte pieces per slot
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s)
        for (int q = lane; q < kChunks; q += 32) {
          cp_async16_ca(reinterpret_cast<unsigned char *>(wtab + (((size_t)st * 2 + 0) * NSLOTS + s) * kNTile) + q * 16,
                        reinterpret_cast<const unsigned char *>(a.expj_tables + (size_t)lead_tix[s] * kNotchN + (size_t)tibi * kNTile) + q * 16);
          if (two)
            cp_async16_ca(reinterpret_cast<unsigned char *>(wtab + (((size_t)st * 2 + 1) * NSLOTS + s) * kNTile) + q * 16,
                          reinterpret_cast<const unsigned char *>(a.expj_tables + (size_t)alt_tix[s] * kNotchN + (size_t)tibi * kNTile) + q * 16);
        }
    } else {
      staged_mask &= ~(1u << st);
    }
    cp_async_commit();
  };

  const float k = a.k, omk = fsub(1.0f, a.k), gain = a.gain;
  const bool unit_gain = (gain == 1.0f);
  const bool out16 = (reinterpret_cast<uintptr_t>(a.out) & 15u) == 0;   // carry in front: only 8-byte aligned
  // One commit group per tile, kNStages - 1 tiles ahead (empty groups past the end keep the count uniform).
  for (int s = 0; s < kNStages - 1; ++s) { if ((uint64_t)s < total_tiles) issue(s); else cp_async_commit(); }
  for (uint64_t tile = 0; tile < total_tiles; ++tile) {
    if (tile + kNStages - 1 < total_tiles) issue(tile + kNStages - 1); else cp_async_commit();
    cp_async_wait<kNStages - 1>();
    __syncwarp();                            // the table tile was copied by other lanes (rows are private)
    const int st = (int)(tile % kNStages);
    const int64_t blk = base + (int64_t)(tile / kTilesPerBlock);
    const int tib = (int)(tile % kTilesPerBlock);
    const bool active = have && blk >= (int64_t)run_begin && blk < (int64_t)own_end;
    const bool write = active && ((uint64_t)blk >= own_begin);
    float2 *outp = a.out + (uint64_t)(active ? blk : 0) * kNotchN + (uint64_t)tib * kNTile;
    unsigned char *orow_base = smem + NotchSmem<NSLOTS>::in + (size_t)warp * 32 * kNPitch;   // the warp's output tile
    if (active) {
      if (tib == 0) {
        // Block start: entry snapshot, epoch switch / resets (sdr.h:97-109).
        while (ep + 1 < a.nepochs && a.epochs[ep + 1].first_block <= (uint64_t)blk) ++ep;
        if ((uint64_t)blk == own_begin && a.seg_entry && !forced_entry) {
#pragma unroll
          for (int s = 0; s < NSLOTS; ++s) a.seg_entry[(size_t)seg * kNotchMaxSlots + s] = make_float2(er[s], ei[s]);
        }
        if (a.epochs[ep].first_block == (uint64_t)blk) {
#pragma unroll
          for (int s = 0; s < NSLOTS; ++s)
            if (a.epochs[ep].reset[s]) { er[s] = 0.f; ei[s] = 0.f; }
        }
      }
      const bool staged = (staged_mask >> st) & 1u;
      const float2 *tab[NSLOTS];       // this lane's table in global memory (used when the tile is not staged)
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) tab[s] = nullptr;
      if (!staged) {
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) tab[s] = a.expj_tables + (size_t)a.epochs[ep].table_index[s] * kNotchN + tib * kNTile;
      }
      const float4 *stab = reinterpret_cast<const float4 *>(wtab + ((size_t)st * 2 + ((set_mask >> st) & 1u)) * NSLOTS * kNTile);   // shared: this lane's set
      float4 *myout = reinterpret_cast<float4 *>(orow_base + (size_t)lane * kNPitch);
      uint32_t lead; { const unsigned char *unused; locate((uint64_t)blk * kNotchN + (uint64_t)tib * kNTile, unused, lead); }
      const unsigned char *myrow = smem + row_off + (size_t)st * 32 * kNPitch;
      // Eight samples per step: loads and the products that do not depend on the
      // estimate are issued first (ILP), then the 8-step serial chain
      // estim = bb*k + estim*(1-k), then the subtraction and the stores.
      constexpr int U = 8;
      // Groups of a tile are unrolled so that the loads / products of group g+1 overlap the
      // serial chain of group g (few warps per scheduler: ILP has to come from the lane itself).
      constexpr int kUnroll = NSLOTS == 1 ? 4 : (NSLOTS == 2 ? 2 : 1);
#pragma unroll kUnroll
      for (int n0 = 0; n0 < kNTile; n0 += U) {
        float2 x[U], e[U][NSLOTS];
        float bkr[U][NSLOTS], bki[U][NSLOTS];
        if (FMT >= 4 && (lead & 1u) == 0) {
          // cf32 rows, 16-byte aligned: two samples per shared-memory load (conflict-free at
          // this row pitch, half the wavefronts of 8-byte loads)
          const float4 *row4 = reinterpret_cast<const float4 *>(myrow) + ((lead + n0) >> 1);
#pragma unroll
          for (int j = 0; j < U; j += 2) {
            const float4 v = row4[j >> 1];
            x[j] = make_float2(v.x, v.y); x[j + 1] = make_float2(v.z, v.w);
            if (FMT == 4 && a.scale != 1.0f) {
              x[j] = make_float2(fmul(x[j].x, a.scale), fmul(x[j].y, a.scale));
              x[j + 1] = make_float2(fmul(x[j + 1].x, a.scale), fmul(x[j + 1].y, a.scale));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < U; ++j) x[j] = row_sample<FMT>(myrow, lead + n0 + j, a.scale);
        }
        if (staged) {                  // warp-uniform: shared-memory broadcasts, two entries per load
#pragma unroll
          for (int s = 0; s < NSLOTS; ++s)
#pragma unroll
            for (int j = 0; j < U; j += 2) {
              const float4 v = stab[(s * kNTile + n0 + j) >> 1];
              e[j][s] = make_float2(v.x, v.y); e[j + 1][s] = make_float2(v.z, v.w);
            }
        } else {
#pragma unroll
          for (int j = 0; j < U; ++j) {
#pragma unroll
            for (int s = 0; s < NSLOTS; ++s) e[j][s] = __ldg(tab[s] + n0 + j);
          }
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
#pragma unroll
          for (int s = 0; s < NSLOTS; ++s) {
            bkr[j][s] = fmul(fadd(fmul(x[j].x, e[j][s].x), fmul(x[j].y, e[j][s].y)), k);
            bki[j][s] = fmul(fadd(fmul(-x[j].x, e[j][s].y), fmul(x[j].y, e[j][s].x)), k);
          }
        float esr[U][NSLOTS], esi[U][NSLOTS];
#pragma unroll
        for (int j = 0; j < U; ++j)
#pragma unroll
          for (int s = 0; s < NSLOTS; ++s) {
            er[s] = fadd(bkr[j][s], fmul(er[s], omk));
            ei[s] = fadd(bki[j][s], fmul(ei[s], omk));
            esr[j][s] = er[s]; esi[j][s] = ei[s];
          }
        if (write) {
          float2 o[U];
#pragma unroll
          for (int j = 0; j < U; ++j) {
            float outr = x[j].x, outi = x[j].y;
#pragma unroll
            for (int s = 0; s < NSLOTS; ++s) {
              outr = fsub(outr, fsub(fmul(esr[j][s], e[j][s].x), fmul(esi[j][s], e[j][s].y)));
              outi = fsub(outi, fadd(fmul(esr[j][s], e[j][s].y), fmul(esi[j][s], e[j][s].x)));
            }
            if (!unit_gain) { outr = fmul(gain, outr); outi = fmul(gain, outi); }
            o[j] = make_float2(outr, outi);
          }
          // into the lane's row of the warp's output tile (conflict-free pitch); the warp writes it out below
#pragma unroll
          for (int j = 0; j < U; j += 2) myout[(n0 + j) >> 1] = make_float4(o[j].x, o[j].y, o[j + 1].x, o[j + 1].y);
        }
      }
    }
    // Output rows leave through the warp as well: 16 lanes per 256-byte row, two rows per store
    // instruction (full 128-byte lines instead of 32 scattered 16-byte pieces).
    const unsigned wmask = __ballot_sync(0xffffffffu, write);
    if (wmask) {
      __syncwarp();
      const uint32_t olo = (uint32_t)reinterpret_cast<uintptr_t>(outp), ohi = (uint32_t)(reinterpret_cast<uintptr_t>(outp) >> 32);
      constexpr int kOutChunks = kNTile * 8 / 16, kOutRows = 32 / kOutChunks;
      const int q = lane % kOutChunks, sub = lane / kOutChunks;
#pragma unroll
      for (int r = 0; r < 32; r += kOutRows) {
        const int row = r + sub;
        const uint32_t lo = __shfl_sync(0xffffffffu, olo, row), hi = __shfl_sync(0xffffffffu, ohi, row);
        if ((wmask >> row) & 1u) {
          const float4 v = *reinterpret_cast<const float4 *>(orow_base + (size_t)row * kNPitch + q * 16);
          unsigned char *dst = reinterpret_cast<unsigned char *>(((uintptr_t)hi << 32) | lo) + q * 16;
          if (out16) st_stream(reinterpret_cast<float4 *>(dst), v);
          else { st_stream(reinterpret_cast<float2 *>(dst), make_float2(v.x, v.y)); st_stream(reinterpret_cast<float2 *>(dst) + 1, make_float2(v.z, v.w)); }
        }
      }
    }
    __syncwarp();
  }
  if (have) {
    if (a.seg_exit) {
#pragma unroll
      for (int s = 0; s < NSLOTS; ++s) a.seg_exit[(size_t)seg * kNotchMaxSlots + s] = make_float2(er[s], ei[s]);
    }
    if (a.seg_exact && !forced_entry) a.seg_exact[seg] = (p.start_kind != 2) ? 1 : 0;
  }
}


template <int FMT, int NSLOTS>
cudaError_t launch_apply_t(const NotchApplyArgs &a, const uint32_t *seg_list, uint32_t nlist,
                           const float2 *guess, cudaStream_t st) {
  static PerDeviceMark configured;   // per device (function attributes belong to the context)
  if (configured.need(1)) {
    cudaError_t e = cudaFuncSetAttribute(k_notch_apply<FMT, NSLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)NotchSmem<NSLOTS>::total);
    if (e != cudaSuccess) return e;
    // Four resident CTAs per SM (4 x 55 KB of rows, table tiles and output tiles) cover the whole
    // grid at the bench size; the e^{j theta} tables are staged in shared memory, so L1 is not needed.
    static const int carve = [] { const char *v = getenv("LDVB_NOTCH_CARVEOUT"); return v ? atoi(v) : 100; }();
    if (carve >= 0 && carve <= 100) {
      e = cudaFuncSetAttribute(k_notch_apply<FMT, NSLOTS>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      if (e != cudaSuccess) return e;
    }
    configured.commit(1);
  }
  const unsigned per_block = kNWarps * 32;
  const uint32_t lanes = seg_list ? nlist : a.nsegs;
  if (!lanes) return cudaSuccess;
  k_notch_apply<FMT, NSLOTS><<<(lanes + per_block - 1) / per_block, per_block, NotchSmem<NSLOTS>::total, st>>>(a, seg_list, nlist, guess);
  return cudaGetLastError();
}

template <int FMT>
cudaError_t launch_apply_f(const NotchApplyArgs &a, const uint32_t *seg_list, uint32_t nlist,
                           const float2 *guess, cudaStream_t st) {
  switch (a.nslots) {
    case 1: return launch_apply_t<FMT, 1>(a, seg_list, nlist, guess, st);
    case 2: return launch_apply_t<FMT, 2>(a, seg_list, nlist, guess, st);
    case 3: return launch_apply_t<FMT, 3>(a, seg_list, nlist, guess, st);
    case 4: return launch_apply_t<FMT, 4>(a, seg_list, nlist, guess, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

namespace {
// entry(j) == exit(j-1), bit for bit, for every segment that started from a guess.
__global__ void k_notch_verify(const float2 *entry, const float2 *exitv, const uint8_t *exact, uint32_t nsegs,
                               int nslots, uint32_t *nfail) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0 || j >= nsegs || exact[j]) return;
  bool same = true;
  for (int s = 0; s < nslots; ++s) {
    const float2 e = entry[(size_t)j * kNotchMaxSlots + s], x = exitv[(size_t)(j - 1) * kNotchMaxSlots + s];
    same = same && __float_as_uint(e.x) == __float_as_uint(x.x) && __float_as_uint(e.y) == __float_as_uint(x.y);
  }
  if (!same) atomicAdd(nfail, 1u);
}
}  // namespace

cudaError_t launch_notch_verify(const NotchApplyArgs &a, uint32_t *nfail, cudaStream_t st) {
  if (a.nsegs < 2) return cudaSuccess;
  k_notch_verify<<<(a.nsegs + 255) / 256, 256, 0, st>>>(a.seg_entry, a.seg_exit, a.seg_exact, a.nsegs, a.nslots, nfail);
  return cudaGetLastError();
}

cudaError_t launch_notch_detect(NotchDetectArgs a, cudaStream_t st) {
  if (a.ndetect <= 0) return cudaSuccess;
  k_notch_detect<<<a.ndetect, 1024, 0, st>>>(a);
  return cudaGetLastError();
}

namespace {
template <int FMT>
void guess_launch_f(const NotchApplyArgs &a, unsigned blocks, uint64_t first, float2 *sums, const float *weights, cudaStream_t st) {
  switch (a.nslots) {
    case 1: k_notch_guess<FMT, 1><<<blocks, 128, 0, st>>>(a, first, sums, weights); break;
    case 2: k_notch_guess<FMT, 2><<<blocks, 128, 0, st>>>(a, first, sums, weights); break;
    case 3: k_notch_guess<FMT, 3><<<blocks, 128, 0, st>>>(a, first, sums, weights); break;
    default: k_notch_guess<FMT, 4><<<blocks, 128, 0, st>>>(a, first, sums, weights); break;
  }
}
}  // namespace

cudaError_t launch_notch_guess(const NotchApplyArgs &a, float2 *sums, const float *weights, cudaStream_t st) {
  // Blocks whose sums can be asked for: from two blocks in front of the first warm-up on.
  const uint64_t lead = (uint64_t)a.warm_blocks + 2;
  const uint64_t first = a.block0 > lead ? a.block0 - lead : 0;
  if (a.nblocks <= first || a.nslots < 1) return cudaSuccess;
  const unsigned blocks = (unsigned)(a.nblocks - first);
  switch (a.fmt) {
    case 0: guess_launch_f<0>(a, blocks, first, sums, weights, st); break;
    case 1: guess_launch_f<1>(a, blocks, first, sums, weights, st); break;
    case 2: guess_launch_f<2>(a, blocks, first, sums, weights, st); break;
    case 3: guess_launch_f<3>(a, blocks, first, sums, weights, st); break;
    case 4: guess_launch_f<4>(a, blocks, first, sums, weights, st); break;
    default: guess_launch_f<5>(a, blocks, first, sums, weights, st); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_notch_apply(NotchApplyArgs a, const uint32_t *seg_list, uint32_t nlist,
                               const float2 *guess, cudaStream_t st) {
  if (a.nsegs == 0 || a.nblocks == 0) return cudaSuccess;
  switch (a.fmt) {
    case 0: return launch_apply_f<0>(a, seg_list, nlist, guess, st);
    case 1: return launch_apply_f<1>(a, seg_list, nlist, guess, st);
    case 2: return launch_apply_f<2>(a, seg_list, nlist, guess, st);
    case 3: return launch_apply_f<3>(a, seg_list, nlist, guess, st);
    case 4: return launch_apply_f<4>(a, seg_list, nlist, guess, st);
    default: return launch_apply_f<5>(a, seg_list, nlist, guess, st);
  }
}

}  // namespace ldvb
