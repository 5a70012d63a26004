// cuda_emu.h -- TEST INFRASTRUCTURE: runs a __global__ function of the product on the host.
//
// The container the kernels are written in has no GPU.  The small control kernels of the receive
// chain (one CTA walking a batch's packet heads, span records or lock mask with block scans, warp
// shuffles and ballots) are integer programs whose whole risk is in their indexing, so their SOURCE
// (`leansdr_b200/csrc/k_ctl_*.cuh`, the very text nvcc compiles) is also compiled by g++ against this
// shim and checked against the kernels they replace on the CPU (`tests/emu/emu_ctl.cpp`,
// `tests/test_ctl_kernels_cpu.py`).  One OS thread per CUDA thread, a sleeping barrier for
// __syncthreads(), a per-warp exchange slot + barrier for shuffles and ballots, GCC __atomic
// builtins for atomics.  Blocks of a grid run one after the other (so `__shared__` can be a
// function-local static).  Nothing here is ever linked into the product.
#pragma once
#include <cuda_runtime.h>   // vector types, cudaError_t (host-side declarations only)

#include <barrier>
#include <condition_variable>
#include <mutex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#define LDVB_DYN_SMEM(name) unsigned char *name = emu::g_dyn_smem

namespace emu {
inline unsigned char *g_dyn_smem = nullptr;   // dynamic shared memory of the block being run (set by the harness)
// Block barrier that SLEEPS (mutex + condition variable): warps that wait for the rest of the block must not spin --
// std::barrier spins and yields first, and with a few dozen waiting host threads per core the one working warp of a
// kernel like k_viterbi crawls.
class SleepBarrier {
 public:
  explicit SleepBarrier(std::ptrdiff_t n) : expected_(n), waiting_(0), gen_(0) {}
  void arrive_and_wait() {
    std::unique_lock<std::mutex> lk(m_);
    if (++waiting_ == expected_) { waiting_ = 0; ++gen_; cv_.notify_all(); return; }
    const unsigned long g = gen_;
    cv_.wait(lk, [&] { return gen_ != g; });
  }
  void arrive_and_drop() {
    std::unique_lock<std::mutex> lk(m_);
    --expected_;
    if (expected_ > 0 && waiting_ == expected_) { waiting_ = 0; ++gen_; cv_.notify_all(); }
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::ptrdiff_t expected_, waiting_;
  unsigned long gen_;
};
struct Warp {
  std::barrier<> bar;
  uint64_t slot[32];
  explicit Warp(int n) : bar(n) {}
};
struct BlockCtx {
  std::unique_ptr<SleepBarrier> bar;
  std::vector<std::unique_ptr<Warp>> warps;
  uint64_t red[3];   // __syncthreads_or / _and / _count
};
inline thread_local BlockCtx *g_blk = nullptr;
inline thread_local Warp *g_warp = nullptr;
inline thread_local int g_lane = 0;
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
constexpr int warpSize = 32;

inline void __syncthreads() { emu::g_blk->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_warp->bar.arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
  if (threadIdx.x == 0) emu::g_blk->red[0] = 0;
  __syncthreads();
  if (pred) __atomic_fetch_or(&emu::g_blk->red[0], 1ull, __ATOMIC_SEQ_CST);
  __syncthreads();
  const int r = (int)emu::g_blk->red[0];
  __syncthreads();
  return r;
}
inline int __syncthreads_and(int pred) {
  if (threadIdx.x == 0) emu::g_blk->red[1] = 1;
  __syncthreads();
  if (!pred) __atomic_store_n(&emu::g_blk->red[1], 0ull, __ATOMIC_SEQ_CST);
  __syncthreads();
  const int r = (int)emu::g_blk->red[1];
  __syncthreads();
  return r;
}

namespace emu {
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
// every lane publishes, reads the lane `src(lane)` (own value when out of range), and leaves together
template <class T, class F> inline T exchange(T v, F src) {
  Warp *w = g_warp;
  w->slot[g_lane] = to_bits(v);
  w->bar.arrive_and_wait();
  const int s = src(g_lane);
  const T r = (s >= 0 && s < 32) ? from_bits<T>(w->slot[s]) : v;
  w->bar.arrive_and_wait();
  return r;
}
}  // namespace emu

template <class T> inline T __shfl_up_sync(unsigned, T v, int o) { return emu::exchange(v, [o](int l) { return l - o; }); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int o) { return emu::exchange(v, [o](int l) { return l + o; }); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int o) { return emu::exchange(v, [o](int l) { return l ^ o; }); }
template <class T> inline T __shfl_sync(unsigned, T v, int s) { return emu::exchange(v, [s](int) { return s & 31; }); }
inline unsigned __ballot_sync(unsigned, int pred) {
  emu::Warp *w = emu::g_warp;
  w->slot[emu::g_lane] = pred ? 1 : 0;
  w->bar.arrive_and_wait();
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) r |= (unsigned)(w->slot[l] & 1) << l;
  w->bar.arrive_and_wait();
  return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  for (int o = 16; o; o >>= 1) { const unsigned u = __shfl_xor_sync(0xffffffffu, v, o); if (u < v) v = u; }
  return v;
}
inline int __reduce_min_sync(unsigned, int v) {
  for (int o = 16; o; o >>= 1) { const int u = __shfl_xor_sync(0xffffffffu, v, o); if (u < v) v = u; }
  return v;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (unsigned)((v << (sh & 31)) >> 32);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (unsigned)(v >> (sh & 31));
}
inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
inline float __uint_as_float(unsigned u) { return emu::from_bits<float>((uint64_t)u); }
template <class T> inline T __ldg(const T *p) { return *p; }

template <class T> inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicMax(T *p, T v) {
  T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return cur;
}
template <class T> inline T atomicMin(T *p, T v) {
  T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (cur > v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return cur;
}

// CUDA's overloaded min/max (mixed signed/unsigned arguments are non-negative wherever the kernels mix them)
template <class A, class B> inline std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)b < (C)a ? (C)b : (C)a; }
template <class A, class B> inline std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)b : (C)a; }

namespace emu {
// Run `body` (a call of the __global__ function) as a grid of 1-D blocks.  A thread that returns
// from the kernel early leaves the barriers (CUDA: exited threads no longer take part).
inline void launch(unsigned grid, unsigned block, const std::function<void()> &body) {
  for (unsigned b = 0; b < grid; ++b) {
    BlockCtx ctx;
    ctx.bar = std::make_unique<SleepBarrier>((std::ptrdiff_t)block);
    const unsigned nwarps = (block + 31) / 32;
    for (unsigned w = 0; w < nwarps; ++w) ctx.warps.push_back(std::make_unique<Warp>((int)std::min(32u, block - 32 * w)));
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t) {
      th.emplace_back([&, t, b] {
        threadIdx = {t, 0, 0};
        blockIdx = {b, 0, 0};
        blockDim = dim3(block, 1, 1);
        gridDim = dim3(grid, 1, 1);
        g_blk = &ctx;
        g_warp = ctx.warps[t / 32].get();
        g_lane = (int)(t & 31);
        body();
        g_warp->bar.arrive_and_drop();
        ctx.bar->arrive_and_drop();
      });
    }
    for (auto &x : th) x.join();
  }
}
}  // namespace emu
