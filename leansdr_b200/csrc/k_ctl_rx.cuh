// k_ctl_rx.cuh -- the seam-resolution kernel of the receiver stage.  Device code only, no launch syntax: compiled by
// nvcc as part of k_rx.cu (inside ldvb's anonymous namespace) AND by g++ against tests/emu/cuda_emu.h, where it is run
// on the host and compared with its predecessor (tests/emu/ctl_v1.cuh) -- see k_ctl_fec.cuh.

// Seam resolution on the device: kept symbol counts, skips, cumulative rotations and
// output offsets of every span (one CTA), so that the host only reads back a few numbers when
// every seam verified.  Every thread owns a contiguous run of spans: a first walk sums its runs'
// counts and rotations, one block-wide scan of the 1024 partial sums gives every run its base, a
// second walk writes the per-span results (round 1 scanned 1024 spans at a time, 54 rounds of
// block scans for one wave of spans: 0.126 ms).
__global__ void __launch_bounds__(1024)
k_rx_plan(const RxSpanInfo *info, const RxSeam *seams, uint32_t nspans, uint32_t span_cap, int nrot,
          int rot0, uint32_t skip0, uint64_t *span_offset, uint32_t *span_skip, uint8_t *span_rot, uint64_t *result /* [9] */) {
  __shared__ unsigned long long s_sum[32];
  __shared__ int s_rot[32];
  __shared__ unsigned int nfail, overflow, nmis, mx_phase, mx_freqw, mx_mu, nfail_loose;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { nfail = 0; overflow = 0; nmis = 0; mx_phase = 0; mx_freqw = 0; mx_mu = 0; nfail_loose = 0; }
  __syncthreads();
  const uint32_t C = (nspans + 1023u) / 1024u;
  const uint32_t j0 = min(nspans, (uint32_t)tid * C), j1 = min(nspans, j0 + C);
  // what span j keeps, drops at its head and how it is rotated against its predecessor
  auto span = [&](uint32_t j, unsigned long long &keep, int &rot, uint32_t &skip, const RxSeam *&before) {
    keep = info[j].n_out;
    before = nullptr;
    if (j > 0) { before = seams + (j - 1); skip = (uint32_t)before->skip_next; rot = before->rot; }
    else { skip = skip0; rot = rot0; }   // seam in front of span 0 (previous rank), 0 otherwise
    keep -= skip;
    if (j + 1 < nspans) keep += (unsigned long long)seams[j].extend_prev;
  };
  unsigned long long ks = 0; int rs = 0;
  {
    unsigned int c_fail = 0, c_over = 0, c_mis = 0, c_loose = 0, m_ph = 0, m_fw = 0, m_mu = 0;
    for (uint32_t j = j0; j < j1; ++j) {
      unsigned long long keep; int rot; uint32_t skip; const RxSeam *sm;
      span(j, keep, rot, skip, sm);
      const RxSpanInfo inf = info[j];
      if (inf.n_out + inf.n_tail > span_cap) ++c_over;
      if (sm) {
        if (!sm->ok) ++c_fail;
        else if (sm->mismatches) ++c_mis;
        if (!sm->ok_loose) ++c_loose;
        if (sm->ok) {   // (non-negative floats order like their bit patterns)
          m_ph = max(m_ph, __float_as_uint(fabsf(sm->dphase)));
          m_fw = max(m_fw, __float_as_uint(fabsf(sm->dfreqw)));
          m_mu = max(m_mu, __float_as_uint(fabsf(sm->dmu)));
        }
      }
      ks += keep; rs += rot;
    }
    if (c_fail) atomicAdd(&nfail, c_fail);
    if (c_over) atomicAdd(&overflow, c_over);
    if (c_mis) atomicAdd(&nmis, c_mis);
    if (c_loose) atomicAdd(&nfail_loose, c_loose);
    if (m_ph) atomicMax(&mx_phase, m_ph);
    if (m_fw) atomicMax(&mx_freqw, m_fw);
    if (m_mu) atomicMax(&mx_mu, m_mu);
  }
  // inclusive scan of (ks, rs) over the block
  unsigned long long kincl = ks; int rincl = rs;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, kincl, o);
    const int b = __shfl_up_sync(0xffffffffu, rincl, o);
    if (lane >= o) { kincl += a; rincl += b; }
  }
  if (lane == 31) { s_sum[warp] = kincl; s_rot[warp] = rincl; }
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
  unsigned long long run = (warp ? s_sum[warp - 1] : 0) + kincl - ks;   // symbols kept in front of this thread's spans
  int rrun = (warp ? s_rot[warp - 1] : 0) + rincl - rs;
  for (uint32_t j = j0; j < j1; ++j) {
    unsigned long long keep; int rot; uint32_t skip; const RxSeam *sm;
    span(j, keep, rot, skip, sm);
    run += keep; rrun += rot;
    span_offset[j + 1] = run;
    span_skip[j] = skip;
    span_rot[j] = (uint8_t)(rrun % nrot);
  }
  if (tid == 0) {
    span_offset[0] = 0;
    result[0] = nfail; result[1] = s_sum[31]; result[2] = (uint64_t)(s_rot[31] % nrot); result[3] = overflow;
    result[4] = nmis; result[5] = mx_phase; result[6] = mx_freqw; result[7] = mx_mu; result[8] = nfail_loose;
  }
}
