// tsan_sanity.cpp -- TEST INFRASTRUCTURE.  Shows that ThreadSanitizer sees, through cuda_emu.h, the two kinds of race a
// CUDA kernel can have in shared memory: across warps without __syncthreads() and inside a warp without __syncwarp()
// (threads of a warp are scheduled independently since Volta).  Usage: tsan_sanity [sync]; without the argument both
// kernels race and ThreadSanitizer must report them, with it they must be clean.
#include "cuda_emu.h"

#include <cstdio>

__global__ void k_block_exchange(int *out, int with_sync) {
  __shared__ int s[64];
  s[threadIdx.x] = threadIdx.x * 3;
  if (with_sync) __syncthreads();
  out[threadIdx.x] = s[(threadIdx.x + 1) & 63];
}
__global__ void k_warp_exchange(int *out, int with_sync) {
  __shared__ int s[32];
  s[threadIdx.x] = threadIdx.x * 3;
  if (with_sync) __syncwarp();
  out[threadIdx.x] = s[(threadIdx.x + 1) & 31];
}
int main(int argc, char **argv) {
  int out[64];
  const int with_sync = argc > 1;
  emu::launch(1, 64, [&] { k_block_exchange(out, with_sync); });
  emu::launch(1, 32, [&] { k_warp_exchange(out, with_sync); });
  printf("done %d\n", out[3]);
  return 0;
}
