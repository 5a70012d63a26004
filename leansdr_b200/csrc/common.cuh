// common.cuh -- shared device helpers for the sm_100a DVB-S kernels.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>

// Dynamic shared memory of a kernel whose text is also compiled on the host by the test shim (tests/emu/cuda_emu.h
// defines its own LDVB_DYN_SMEM).
#ifndef LDVB_DYN_SMEM
#define LDVB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace ldvb {

// ---------------------------------------------------------- per-device opt-ins
// cudaFuncSetAttribute is per device (context): a process may hold handles on several GPUs,
// possibly on different threads, so "already configured" is remembered per device ordinal.
// need(v) returns true when this device has not been configured for at least `v` yet (the caller then
// sets the attribute and calls commit(v)); racing threads at worst set the same attribute twice.
struct PerDeviceMark {
  static constexpr int kMaxDev = 64;
  std::atomic<size_t> mark[kMaxDev];
  PerDeviceMark() { for (auto &m : mark) m.store(0); }
  int dev() const { int d = 0; if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDev) d = kMaxDev - 1; return d; }
  bool need(size_t v) const { return mark[dev()].load(std::memory_order_acquire) < v; }
  void commit(size_t v) { auto &m = mark[dev()]; size_t cur = m.load(); while (cur < v && !m.compare_exchange_weak(cur, v)) {} }
};

// ----------------------------------------------------------------- arithmetic
// Every float operation on the bit-exact path goes through the _rn intrinsics:
// they are never contracted into FMAs, so the device rounds each product and
// sum exactly like the reference's scalar SSE2 build (hard part 2).
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// complex<float> operator* of the reference (math.h:38-41):
//   (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fsub(fmul(a.x, b.x), fmul(a.y, b.y)),
                     fadd(fmul(a.x, b.y), fmul(a.y, b.x)));
}

// (T)value conversions of the reference truncate toward zero (cvttss2si).
__device__ __forceinline__ int f2i_trunc(float a) { return __float2int_rz(a); }

// ------------------------------------------------------------ TMA bulk copies
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier
// (SASS: UBLKCP).  dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 16-byte asynchronous copies global -> shared (SASS LDGSTS), L1-bypassing.
// Used where each lane of a warp walks its own place in a stream: rows of a few
// hundred bytes are too small for the TMA engine (measured: ~0.6 us per 272-byte
// cp.async.bulk request, serialised per SM), so the 32 rows of a warp are
// fetched cooperatively, 16 B per lane per instruction, contiguous within a row.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// Same, allocating in L1 (.ca): the two 16-byte halves of a 32-byte sector then cost one L2
// request instead of two (the second one merges with the outstanding miss).
__device__ __forceinline__ void cp_async16_ca(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Streaming stores: results are written once and read by a later kernel.
__device__ __forceinline__ void st_stream(float2 *p, float2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { __stcs(p, v); }

}  // namespace ldvb
