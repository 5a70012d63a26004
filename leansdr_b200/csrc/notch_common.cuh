// notch_common.cuh -- device helpers shared by the auto_notch kernels (k_notch.cu, k_notchfir.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ldvb {
namespace {

template <int FMT>
__device__ __forceinline__ float2 ld_raw(const RawSrc &src, uint64_t idx, float scale) {
  const void *raw = src.head;
  if (src.main && idx >= src.c0) { raw = src.main; idx -= src.c0; }
  if (FMT == 0) { uchar2 v = reinterpret_cast<const uchar2 *>(raw)[idx];
    return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128)); }
  if (FMT == 1) { char2 v = reinterpret_cast<const char2 *>(raw)[idx];
    return make_float2((float)(int)v.x, (float)(int)v.y); }
  if (FMT == 2) { ushort2 v = reinterpret_cast<const ushort2 *>(raw)[idx];
    return make_float2((float)((int)v.x - 32768), (float)((int)v.y - 32768)); }
  if (FMT == 3) { short2 v = reinterpret_cast<const short2 *>(raw)[idx];
    return make_float2((float)(int)v.x, (float)(int)v.y); }
  float2 v = __ldg(reinterpret_cast<const float2 *>(raw) + idx);
  if (FMT == 4) v = make_float2(fmul(v.x, scale), fmul(v.y, scale));
  return v;
}

// Segment geometry shared by the guess and apply kernels.
struct SegPlan {
  uint64_t own_begin, own_end, run_begin;  // blocks
  int epoch;                               // epoch of run_begin
  int start_kind;                          // 0 exact carried state, 1 exact zero (full reset), 2 guess
};

__device__ __forceinline__ SegPlan plan_segment(const NotchApplyArgs &a, uint32_t seg) {
  SegPlan p;
  p.own_begin = a.block0 + (uint64_t)seg * a.seg_blocks;
  p.own_end = p.own_begin + a.seg_blocks;
  if (p.own_end > a.nblocks) p.own_end = a.nblocks;
  int ep = 0;
  while (ep + 1 < a.nepochs && a.epochs[ep + 1].first_block <= p.own_begin) ++ep;
  p.epoch = ep;
  if (seg == 0 && a.first_exact) { p.run_begin = p.own_begin; p.start_kind = 0; return p; }
  // One warm-up block, never across an epoch start (tables / resets change there).
  uint64_t wb = (p.own_begin > a.warm_blocks) ? p.own_begin - a.warm_blocks : 0;
  if (wb < a.epochs[ep].first_block) wb = a.epochs[ep].first_block;
  p.run_begin = wb;
  if (wb == 0 && ep == 0 && a.first_exact) { p.start_kind = 0; return p; }   // reaches the carried state
  if (a.epochs[ep].first_block == wb) {
    bool all = true;
    for (int s = 0; s < a.nslots; ++s) all = all && (a.epochs[ep].reset[s] != 0);
    if (all) { p.start_kind = 1; return p; }
  }
  p.start_kind = 2;
  return p;
}

// Start state of a segment whose exact run begins at block p.run_begin (start_kind 2): the
// estimate forgets with (1-k)^n, so the two blocks in front of it (8192 samples, weight of
// anything older < 1e-7) decide it.  History never reaches across the start of the epoch
// (tables change there); at the very start of the stream it runs into the carried state.
__device__ __forceinline__ float2 guess_from_sums(const NotchApplyArgs &a, const SegPlan &p, const float2 *sums, int s) {
  const uint64_t floor_b = a.epochs[p.epoch].first_block;
  const uint64_t rb = p.run_begin;
  const uint64_t nh = (rb - floor_b) < 2 ? (rb - floor_b) : 2;     // history blocks available
  const float w4096 = a.w_block;                                     // (1-k)^4096
  float2 g = make_float2(0.f, 0.f);
  if (nh >= 1) g = sums[(rb - 1) * kNotchMaxSlots + s];
  if (nh >= 2) { const float2 o = sums[(rb - 2) * kNotchMaxSlots + s]; g.x += o.x * w4096; g.y += o.y * w4096; }
  if (nh < 2 && !a.epochs[p.epoch].reset[s] && p.epoch == 0 && floor_b == 0 && a.first_exact) {
    const float w = nh ? w4096 : 1.0f;
    g.x += a.state_in->slot[s].est_re * w;
    g.y += a.state_in->slot[s].est_im * w;
  }
  return g;
}

template <int FMT>
__device__ __forceinline__ float2 row_sample(const unsigned char *row, uint32_t idx, float scale) {
  if (FMT == 0) { uchar2 v = reinterpret_cast<const uchar2 *>(row)[idx];
    return make_float2((float)((int)v.x - 128), (float)((int)v.y - 128)); }
  if (FMT == 1) { char2 v = reinterpret_cast<const char2 *>(row)[idx];
    return make_float2((float)(int)v.x, (float)(int)v.y); }
  if (FMT == 2) { ushort2 v = reinterpret_cast<const ushort2 *>(row)[idx];
    return make_float2((float)((int)v.x - 32768), (float)((int)v.y - 32768)); }
  if (FMT == 3) { short2 v = reinterpret_cast<const short2 *>(row)[idx];
    return make_float2((float)(int)v.x, (float)(int)v.y); }
  float2 v = reinterpret_cast<const float2 *>(row)[idx];
  if (FMT == 4 && scale != 1.0f) v = make_float2(fmul(v.x, scale), fmul(v.y, scale));  // x*1 == x
  return v;
}


}  // namespace
}  // namespace ldvb
