// ctl_v1.cuh -- TEST INFRASTRUCTURE.  The control kernels exactly as they ran (and were checked against the oracle,
// 144 GPU tests) on B200 at commit f37a100, kept as the reference the rewritten kernels of
// leansdr_b200/csrc/k_ctl_*.cuh are compared with under the host emulation (tests/emu/cuda_emu.h).
// Included inside namespace ldvb::v1.
// ---- k_rx.cu @ f37a100
// Seam resolution on the device: kept symbol counts, skips, cumulative rotations and
// output offsets of every span (one CTA, block-wide scans), so that the host only
// reads back two numbers when every seam verified.
__global__ void __launch_bounds__(1024)
k_rx_plan(const RxSpanInfo *info, const RxSeam *seams, uint32_t nspans, uint32_t span_cap, int nrot,
          int rot0, uint32_t skip0, uint64_t *span_offset, uint32_t *span_skip, uint8_t *span_rot, uint64_t *result /* [4] */) {
  __shared__ unsigned long long s_sum[32];
  __shared__ int s_rot[32];
  __shared__ unsigned long long carry_sum;
  __shared__ int carry_rot;
  __shared__ unsigned int nfail, overflow, nmis, mx_phase, mx_freqw, mx_mu, nfail_loose;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_sum = 0; carry_rot = 0; nfail = 0; overflow = 0; nmis = 0; mx_phase = 0; mx_freqw = 0; mx_mu = 0; nfail_loose = 0; }
  __syncthreads();
  for (uint32_t base = 0; base < nspans; base += 1024) {
    const uint32_t j = base + tid;
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
    // inclusive scans (sum of keep, sum of rot mod nrot) across the block
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
    const unsigned long long wbase = warp ? s_sum[warp - 1] : 0;
    const int wrot = warp ? s_rot[warp - 1] : 0;
    const unsigned long long incl = carry_sum + wbase + ks;
    const int rincl = (carry_rot + wrot + rs) % nrot;
    if (j < nspans) {
      span_offset[j + 1] = incl;
      span_skip[j] = skip;
      span_rot[j] = (uint8_t)rincl;
    }
    __syncthreads();
    if (tid == 1023) { carry_sum = incl; carry_rot = rincl; }
    __syncthreads();
  }
  if (tid == 0) {
    span_offset[0] = 0;
    result[0] = nfail; result[1] = carry_sum; result[2] = (uint64_t)carry_rot; result[3] = overflow;
    result[4] = nmis; result[5] = mx_phase; result[6] = mx_freqw; result[7] = mx_mu; result[8] = nfail_loose;
  }
}

// ---- k_fec.cu @ f37a100
__device__ __forceinline__ unsigned par64(uint64_t v) { return __popcll(v) & 1; }
// of every symbol it needs (plus the 32 carried ones) into a shared-memory bit string
// -- 16 symbols per word, loaded 4 symbols per 16-byte access, each symbol read from
// HBM exactly once -- then every thread extracts the 64-bit register of a bit group
// with three shared loads and two funnel shifts.
constexpr int kDcBytes = 1024;
constexpr int kDcWords = 576;

__global__ void __launch_bounds__(256)
k_deconv_tiled(DeconvArgs a, uint64_t nsym) {
  __shared__ uint32_t s_bits[kDcWords];
  const int pp = a.punctperiod, half = a.punctweight / 2;
  const int64_t k0 = (a.n_in >= 64) ? 0 : (64 - a.n_in) / 2;
  const uint64_t b0 = (uint64_t)blockIdx.x * kDcBytes;
  if (b0 >= a.nbytes) return;
  const uint32_t nb = (uint32_t)min((uint64_t)kDcBytes, a.nbytes - b0);
  // Bit groups touched by this CTA's bytes (stream bit i >= n_out belongs to group (i-n_out)/pp).
  const int64_t bit_first = (int64_t)8 * b0, bit_last = (int64_t)8 * (b0 + nb) - 1;
  const int64_t g0 = (bit_first > a.n_out) ? (bit_first - a.n_out) / pp : 0;
  const int64_t g1 = (bit_last >= a.n_out) ? (bit_last - a.n_out) / pp : -1;
  // Extended symbol stream E: E[0..31] = the carried register, E[32+s] = symbol s.
  // The register of group g is E[K_g .. K_g+32), K_g = k0 + g*half.
  // e_base: at or below the first register, and such that symbols + (e_base - 32) is 16-byte aligned whatever the
  // alignment of `symbols` itself (the stream's read position advances by arbitrary symbol counts).
  const int mis = (int)((reinterpret_cast<uintptr_t>(a.symbols) >> 2) & 3u);
  const int64_t e_base = ((k0 + g0 * half + mis) & ~(int64_t)15) - mis;
  const int64_t e_end = (g1 >= 0) ? k0 + g1 * half + 32 : e_base;
  const int nwords = (int)((e_end - e_base + 15) / 16) + 2;
  for (int w = threadIdx.x; w < nwords && w < kDcWords; w += blockDim.x) {
    const int64_t e0 = e_base + (int64_t)16 * w;
    uint32_t word = 0;
    if (e0 >= 32 && (uint64_t)(e0 - 32 + 16) <= nsym) {
      const uint4 *src = reinterpret_cast<const uint4 *>(a.symbols + (e0 - 32));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 v = __ldg(src + q);
        word |= (uint32_t)a.hyp[(v.x >> 16) & 3u] << (30 - 8 * q);
        word |= (uint32_t)a.hyp[(v.y >> 16) & 3u] << (28 - 8 * q);
        word |= (uint32_t)a.hyp[(v.z >> 16) & 3u] << (26 - 8 * q);
        word |= (uint32_t)a.hyp[(v.w >> 16) & 3u] << (24 - 8 * q);
      }
    } else {
      for (int i = 0; i < 16; ++i) {
        const int64_t e = e0 + i;
        uint32_t code = 0;
        if (e < 0) code = 0;                       // (in front of the carried register: never part of a group)
        else if (e < 32) code = (uint32_t)(a.reg_in >> (2 * (31 - e))) & 3u;
        else if ((uint64_t)(e - 32) < nsym) code = a.hyp[(a.symbols[e - 32] >> 16) & 3u];
        word |= code << (30 - 2 * i);
      }
    }
    s_bits[w] = word;
  }
  __syncthreads();
  auto reg_at = [&](int64_t K) -> uint64_t {   // E[K .. K+32) as a 64-bit string
    const int64_t bo = 2 * (K - e_base);
    const int wi = (int)(bo >> 5), sh = (int)(bo & 31);
    const uint32_t hi = s_bits[wi], mid = s_bits[wi + 1], lo = s_bits[wi + 2];
    const uint32_t r_hi = __funnelshift_l(mid, hi, sh), r_lo = __funnelshift_l(lo, mid, sh);
    return ((uint64_t)r_hi << 32) | r_lo;
  };
  if (a.err_out) {
    // readerrors (dvb.h:391-412): every group whose first bit lies in this CTA's bytes, all pp bits of it.
    unsigned err = 0;
    for (uint32_t t = threadIdx.x; t < nb; t += blockDim.x) {
      const int64_t lo = max((int64_t)8 * (b0 + t), (int64_t)a.n_out), hi = (int64_t)8 * (b0 + t) + 8;
      for (int64_t g = (lo - a.n_out + pp - 1) / pp; a.n_out + g * pp < hi; ++g) {
        const uint64_t reg = reg_at(k0 + g * half);
        for (int b = pp - 1; b >= 0; --b) err += par64(reg & a.deconv[b]) ^ par64(reg & a.deconv2[b]);
      }
    }
    for (int o = 16; o; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
    if ((threadIdx.x & 31) == 0 && err) atomicAdd(a.err_out, (unsigned long long)err);
    return;
  }
  for (uint32_t t = threadIdx.x; t < nb; t += blockDim.x) {
    const uint64_t j = b0 + t;
    unsigned byte = 0;
    int64_t bit = (int64_t)8 * j;
    int got = 0;
    while (got < 8 && bit < a.n_out) {   // bits still held by the carried accumulator
      byte = (byte << 1) | (unsigned)((a.out_acc >> (a.n_out - 1 - bit)) & 1);
      ++bit; ++got;
    }
    if (got < 8) {
      int64_t g = (bit - a.n_out) / pp;
      int within = (int)((bit - a.n_out) % pp);
      while (got < 8) {
        const uint64_t reg = reg_at(k0 + g * half);
        for (int b = pp - 1 - within; b >= 0 && got < 8; --b) {
          byte = (byte << 1) | par64(reg & a.deconv[b]);
          ++got;
        }
        within = 0;
        ++g;
      }
    }
    a.out[j] = (uint8_t)byte;
  }
}

__device__ int sync_search_window(const uint8_t *bytes, uint64_t pos, SyncState &st, int *s_best,
                                  int *s_pol, int *s_ph) {
  const int i = threadIdx.x;
  if (i == 0) *s_best = 1 << 30;
  __syncthreads();
  int pol = 0, ph8 = -1;
  bool hit = false;
  if (i < 204) {
    int np = 0, nn = 0, ph_p = -1, ph_n = -1;
    for (int j = 0; j < 8; ++j) {
      const unsigned b = (((unsigned)bytes[pos + i + 204 * j] << 8 | bytes[pos + i + 204 * j + 1]) >> st.bitphase) & 0xffu;
      if (b == 0x47u) { ++np; ph_n = (8 - j) & 7; }
      if (b == 0xb8u) { ++nn; ph_p = (8 - j) & 7; }
    }
    int nsyncs;
    if (np > nn) { pol = 0; nsyncs = np; ph8 = ph_p; }
    else { pol = 0xff; nsyncs = nn; ph8 = ph_n; }
    hit = (nsyncs >= 4 && ph8 >= 0);
    if (hit) atomicMin(s_best, i);
    if (i == 203) { s_pol[1] = pol; s_ph[1] = ph8; }   // what a fruitless scan leaves behind
  }
  __syncthreads();
  const int best = *s_best;
  if (hit && i == best) { s_pol[0] = pol; s_ph[0] = ph8; }
  __syncthreads();
  if (best < 204) {
    st.polarity = s_pol[0]; st.phase8 = s_ph[0];
    int skip = best;
    if (!best) { skip = 204; st.phase8 = (st.phase8 + 1) & 7; }
    st.synchronized = 1;
    st.lock_timeleft = 4;
    st.locktime = 0;
    return skip;
  }
  st.polarity = s_pol[1]; st.phase8 = s_ph[1];
  return 0;
}

__global__ void __launch_bounds__(256)
k_sync_track(const uint8_t *bytes, uint64_t nbytes, const SyncState *st_in,
             const uint32_t *bad_words, uint64_t npackets_flagged, SyncResult *res) {
  __shared__ int s_best, s_pol[2], s_ph[2];
  SyncState st = *st_in;          // every thread keeps an identical copy
  SyncResult r;
  r.consumed = 0; r.produced = 0; r.need_next_sync = 0; r.events = 0;
  auto event = [&](int v, uint64_t pos) {
    if (r.events < 16) { r.event_val[r.events] = v; r.event_pos[r.events] = pos; }
    ++r.events;
  };
  if (st.report_state) { event(0, 0); st.report_state = 0; }
  if (st.synchronized) {
    if (threadIdx.x != 0) return;
    // run_decoding (dvb.h:842-874): walk the mask until the lock times out.
    uint64_t p = 0;
    bool unlocked = false;
    while (p < npackets_flagged) {
      const uint32_t w = bad_words[p >> 5];
      const uint64_t lim = min(npackets_flagged, (p & ~(uint64_t)31) + 32);
      if (w == 0 && (p & 31) == 0 && lim - p == 32) {  // 32 good packets
        st.lock_timeleft = 3;
        st.locktime += 32;
        p += 32;
        continue;
      }
      for (; p < lim; ++p) {
        ++st.locktime;
        if (!((w >> (p & 31)) & 1)) st.lock_timeleft = 4;
        --st.lock_timeleft;
        if (!st.lock_timeleft) { unlocked = true; ++p; break; }
      }
      if (unlocked) break;
    }
    st.phase8 = (int)((st.phase8 + p) & 7);
    r.consumed = 204 * p;
    r.produced = 204 * p;
    if (unlocked) {
      st.synchronized = 0;
      st.next_sync_count = 0;
      event(0, r.consumed);
    }
  } else {
    // run_searching (dvb.h:755-779): one bit phase per 8-packet window; a full
    // sweep of the 8 phases without lock counts towards next_sync().  The sweep
    // counter advances once per wrap (the reference's default buffering never
    // sees two wraps inside one run() call).
    uint64_t pos = 0;
    const uint64_t chunk = 204 * 8;
    if (st.fastlock) {
      // run_searching_fast (dvb.h:781-796): at every resync_period-th packet position all eight
      // bit phases are tried in order; the position advances by ONE packet.
      bool locked = false;
      while (nbytes - pos >= chunk + 1) {
        if (st.resync_phase == 0) {
          for (st.bitphase = 0; st.bitphase <= 7; ++st.bitphase) {
            const int skip = sync_search_window(bytes, pos, st, &s_best, s_pol, s_ph);
            if (skip) { pos += skip; event(1, pos); locked = true; break; }
          }
          if (locked) break;
        }
        pos += 204;
        if (++st.resync_phase >= st.resync_period) st.resync_phase = 0;
      }
      r.consumed = pos;
      if (threadIdx.x != 0) return;
      r.st = st;
      *res = r;
      return;
    }
    while (nbytes - pos >= chunk + 1) {
      const int skip = sync_search_window(bytes, pos, st, &s_best, s_pol, s_ph);
      if (skip) {
        pos += skip;
        event(1, pos);
        break;
      }
      pos += chunk;
      if (++st.bitphase == 8) {
        st.bitphase = 0;
        if (++st.next_sync_count >= 3) {
          st.next_sync_count = 0;
          r.need_next_sync = 1;
          break;
        }
      }
    }
    r.consumed = pos;
    if (threadIdx.x != 0) return;
  }
  r.st = st;
  *res = r;
}

// ============================================================ de-randomiser
// derandomizer::run (dvb.h:1131-1158).  pos_p = 188*((p - r_p) mod 8) with r_p
// the last packet <= p whose first byte is an inverted sync (0xB8 or 0xB8^0x55);
// before the first reset the carried position keeps cycling.  One CTA scans the
// packet heads in tiles (inclusive max-scan of reset indices + exclusive sum of
// kept packets); a second kernel XORs and writes the kept packets.
__global__ void __launch_bounds__(1024)
k_derand_scan(DerandArgs a) {
  __shared__ long long s_last[32];
  __shared__ unsigned s_cnt[32];
  __shared__ long long carry_last;   // last reset index so far, or -1 - pos_in/188 sentinel
  __shared__ unsigned long long carry_kept, carry_errs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_last = -1; carry_kept = 0; carry_errs = 0; }
  __syncthreads();
  const long long start_phase = a.pos_in / 188;  // packets since the (virtual) last reset
  for (uint64_t base = 0; base < a.npackets; base += 1024) {
    const uint64_t p = base + tid;
    const bool valid = p < a.npackets;
    unsigned head = 0;
    if (valid) head = a.rts[188 * p];
    const bool reset = valid && (head == 0xb8u || head == (0xb8u ^ 0x55u));
    long long last = reset ? (long long)p : -1;
    // inclusive max-scan inside the warp
    for (int o = 1; o < 32; o <<= 1) {
      long long v = __shfl_up_sync(0xffffffffu, last, o);
      if (lane >= o && v > last) last = v;
    }
    if (lane == 31) s_last[warp] = last;
    __syncthreads();
    if (warp == 0) {
      long long v = s_last[lane];
      for (int o = 1; o < 32; o <<= 1) {
        long long u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o && u > v) v = u;
      }
      s_last[lane] = v;
    }
    __syncthreads();
    if (warp > 0 && s_last[warp - 1] > last) last = s_last[warp - 1];
    if (carry_last > last) last = carry_last;
    int pos;
    if (last >= 0) pos = (int)(((long long)p - last) & 7) * 188;
    else pos = (int)(((long long)p + start_phase) & 7) * 188;
    bool keep = false;
    if (valid) keep = ((head ^ a.pattern[pos]) == 0x47u);
    int nerr = (valid && a.flags) ? a.flags[2 * p + 1] : 0;
    // exclusive sum of kept packets
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const unsigned before = __popc(bal & ((1u << lane) - 1));
    for (int o = 16; o; o >>= 1) nerr += __shfl_xor_sync(0xffffffffu, nerr, o);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    unsigned wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) { if (w < warp) wbase += s_cnt[w]; total += s_cnt[w]; }
    if (valid) {
      // scratch[p] = output index (bit 31 set when dropped), low bits of pos in scratch2
      a.scratch[p] = keep ? (unsigned)(carry_kept + wbase + before) : 0xffffffffu;
      a.scratch[a.npackets + p] = (unsigned)pos;
    }
    __syncthreads();
    if (lane == 0 && nerr) atomicAdd(&carry_errs, (unsigned long long)nerr);
    if (tid == 1023) {
      carry_last = last;
    }
    if (tid == 0) carry_kept += total;
    __syncthreads();
  }
  if (tid == 0) {
    long long last = carry_last;
    int pos_out;
    if (last >= 0) pos_out = (int)(((long long)a.npackets - last) & 7) * 188;
    else pos_out = (int)(((long long)a.npackets + start_phase) & 7) * 188;
    a.counts[0] = carry_kept;
    a.counts[1] = a.npackets - carry_kept;
    a.counts[2] = (uint64_t)pos_out;
    a.counts[3] = carry_errs;
  }
}
