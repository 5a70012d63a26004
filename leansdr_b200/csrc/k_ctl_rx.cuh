// k_ctl_rx.cuh -- the seam-resolution kernels of the receiver stage.  Device code only, no launch syntax: compiled by
// nvcc as part of k_rx.cu (inside ldvb's anonymous namespace) AND by g++ against tests/emu/cuda_emu.h, where it is run
// on the host and compared with its predecessor (tests/emu/ctl_v1.cuh) -- see k_ctl_fec.cuh.

// Seam resolution on the device: kept symbol counts, skips, cumulative rotations and output offsets of every span, so
// that the host only reads back a few numbers when every seam verified.  Two grids of 1024-span CTAs:
//   k_rx_plan_local  coalesced loads of the span / seam records (consecutive threads, consecutive spans), one block
//                    scan, LOCAL offsets and rotations, the CTA's totals, the statistics of the seams;
//   k_rx_plan_apply  every CTA sums the totals of the CTAs in front of it and rebases its spans.
// History: round 1 walked the batch in one CTA, 1024 spans per round (54 rounds of loads + two barriers: 0.126 ms at
// 55 704 spans); a single scan over per-thread contiguous runs was slower still (0.35 ms measured on B200: every lane
// of a load instruction in another 128-byte line).  `totals` = [nblocks] pairs {sum of keep, sum of rot}.
__global__ void __launch_bounds__(1024)
k_rx_plan_local(const RxSpanInfo *info, const RxSeam *seams, uint32_t nspans, uint32_t span_cap, int nrot,
                int rot0, uint32_t skip0, uint64_t *span_offset, uint32_t *span_skip, uint8_t *span_rot,
                unsigned long long *totals, unsigned long long *result /* [9], zeroed */) {
  __shared__ unsigned long long s_sum[32];
  __shared__ int s_rot[32];
  __shared__ unsigned int nfail, overflow, nmis, mx_phase, mx_freqw, mx_mu, nfail_loose;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { nfail = 0; overflow = 0; nmis = 0; mx_phase = 0; mx_freqw = 0; mx_mu = 0; nfail_loose = 0; }
  __syncthreads();
  const uint32_t j = blockIdx.x * blockDim.x + (uint32_t)tid;   // (1024 threads in the library; any whole number of warps)
  unsigned long long keep = 0; int rot = 0; uint32_t skip = 0;
  if (j < nspans) {
    const RxSpanInfo inf = info[j];
    if (inf.n_out + inf.n_tail > span_cap) atomicAdd(&overflow, 1u);
    keep = inf.n_out;
    if (j > 0) {
      const RxSeam sm = seams[j - 1];
      if (!sm.ok) atomicAdd(&nfail, 1u);
      else if (sm.mismatches) atomicAdd(&nmis, 1u);
      if (!sm.ok_loose) atomicAdd(&nfail_loose, 1u);
      if (sm.ok) {   // (non-negative floats order like their bit patterns)
        atomicMax(&mx_phase, __float_as_uint(fabsf(sm.dphase)));
        atomicMax(&mx_freqw, __float_as_uint(fabsf(sm.dfreqw)));
        atomicMax(&mx_mu, __float_as_uint(fabsf(sm.dmu)));
      }
      skip = (uint32_t)sm.skip_next; rot = sm.rot;
    } else {
      skip = skip0; rot = rot0;   // seam in front of span 0 (previous rank), 0 otherwise
    }
    keep -= skip;
    if (j + 1 < nspans) keep += (unsigned long long)seams[j].extend_prev;
  }
  // inclusive scans (sum of keep, sum of rot) across the block
  unsigned long long ks = keep; int rs = rot;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, ks, o);
    const int b = __shfl_up_sync(0xffffffffu, rs, o);
    if (lane >= o) { ks += a; rs += b; }
  }
  if (lane == 31) { s_sum[warp] = ks; s_rot[warp] = rs; }
  __syncthreads();
  if (warp == 0) {
    unsigned long long a = s_sum[lane]; int b = s_rot[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long a2 = __shfl_up_sync(0xffffffffu, a, o);
      const int b2 = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += a2; b += b2; }
    }
    s_sum[lane] = a; s_rot[lane] = b;
  }
  __syncthreads();
  const unsigned long long incl = (warp ? s_sum[warp - 1] : 0) + ks;
  const int rincl = (warp ? s_rot[warp - 1] : 0) + rs;
  if (j < nspans) {
    span_offset[j + 1] = incl;                 // local: k_rx_plan_apply adds the CTAs in front
    span_skip[j] = skip;
    span_rot[j] = (uint8_t)(rincl % nrot);
  }
  if (tid == (int)blockDim.x - 1) { totals[2 * blockIdx.x] = incl; totals[2 * blockIdx.x + 1] = (unsigned long long)rincl; }
  if (tid == 0) {
    if (nfail) atomicAdd(result + 0, (unsigned long long)nfail);
    if (overflow) atomicAdd(result + 3, (unsigned long long)overflow);
    if (nmis) atomicAdd(result + 4, (unsigned long long)nmis);
    if (mx_phase) atomicMax(result + 5, (unsigned long long)mx_phase);
    if (mx_freqw) atomicMax(result + 6, (unsigned long long)mx_freqw);
    if (mx_mu) atomicMax(result + 7, (unsigned long long)mx_mu);
    if (nfail_loose) atomicAdd(result + 8, (unsigned long long)nfail_loose);
  }
}

__global__ void __launch_bounds__(1024)
k_rx_plan_apply(uint32_t nspans, int nrot, uint64_t *span_offset, uint8_t *span_rot, const unsigned long long *totals,
                unsigned long long *result) {
  __shared__ unsigned long long s_base, s_rbase;
  const int tid = threadIdx.x;
  if (tid < 32) {   // the totals of the CTAs in front of this one
    unsigned long long a = 0, r = 0;
    for (uint32_t b = (uint32_t)tid; b < blockIdx.x; b += 32) { a += totals[2 * b]; r += totals[2 * b + 1]; }
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); r += __shfl_xor_sync(0xffffffffu, r, o); }
    if (tid == 0) { s_base = a; s_rbase = r; }
  }
  __syncthreads();
  const unsigned long long base = s_base, rbase = s_rbase;
  const uint32_t j = blockIdx.x * blockDim.x + (uint32_t)tid;   // (1024 threads in the library; any whole number of warps)
  if (j < nspans) {
    span_offset[j + 1] += base;
    span_rot[j] = (uint8_t)(((unsigned long long)span_rot[j] + rbase) % (unsigned long long)nrot);
  }
  if (tid == 0 && blockIdx.x == 0) span_offset[0] = 0;
  if (tid == 0 && blockIdx.x == gridDim.x - 1) {
    result[1] = base + totals[2 * blockIdx.x];
    result[2] = (rbase + totals[2 * blockIdx.x + 1]) % (unsigned long long)nrot;
  }
}
