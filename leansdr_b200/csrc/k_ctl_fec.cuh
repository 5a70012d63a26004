// k_ctl_fec.cuh -- the control kernels of the byte stages (deconvolution tiles, MPEG sync tracking, de-randomiser
// scan).  Device code only, no launch syntax: this text is compiled by nvcc as part of k_fec.cu (included inside
// ldvb's anonymous namespace) AND by g++ against tests/emu/cuda_emu.h, where every kernel is run on the host and
// compared bit for bit with its predecessor (tests/emu/ctl_v1.cuh, the kernels as verified on B200) -- the way these
// integer programs are checked when no GPU is at hand.
//
// What they have in common: each one used to be a single CTA (or thread) walking the batch in 1024-element rounds,
// one global-load latency and several block barriers per round (k_derand_scan 0.18 ms, k_sync_track 0.19 ms of a
// 4.4 ms step for a few hundred KB of data).  Here the lock tracker stages its mask in shared memory and takes the
// all-good case as a reduction (0.19 -> 0.01 ms on B200), and the de-randomiser scan is spread over tiles of 1024
// packets on as many CTAs with a one-CTA chain step in between.

// =============================================================== deconvolution
// deconvol_sync::readbyte (dvb.h:369-389) as a position-indexed computation.
// Bit group g (punctperiod bits) is the parity of the 64-bit IQ shift register
// after K_g = k0 + g*(punctweight/2) symbols, k0 = symbols needed to fill the
// carried register to 64 bits.  The register holds the last 32 symbols, two IQ
// bits each, newest in the LSBs.  Output byte j = stream bits [8j, 8j+8) where
// the carried accumulator supplies the first n_out bits.
//
// Tiled: a CTA produces 1024 output bytes.  It first packs the IQ bit pairs of every symbol it needs (plus the 32
// carried ones) into a shared-memory bit string, 16 symbols per word.  Symbols are fetched 4 per lane with one
// 16-byte load, consecutive lanes on consecutive addresses (a warp instruction covers 512 contiguous bytes; round 2
// had every lane walk its own 64 bytes: four times the L1 wavefronts), the hypothesis map lives in a register, and
// four lanes merge their bit pairs into a word with two shuffles.  Every thread then extracts the 64-bit register of
// a bit group from a 128-bit window held in registers (one window per output byte at rate 1/2, where the eight
// groups of a byte are 2 bits apart).
constexpr int kDcBytes = 1024;
constexpr int kDcWords = 576;

__global__ void __launch_bounds__(256)
k_deconv_tiled(DeconvArgs a, uint64_t nsym) {
  __shared__ uint32_t s_bits[kDcWords + 4];
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
  const int nwords = min((int)((e_end - e_base + 15) / 16) + 2, kDcWords);
  const uint32_t hyp_lut = (uint32_t)a.hyp[0] | (uint32_t)a.hyp[1] << 2 | (uint32_t)a.hyp[2] << 4 | (uint32_t)a.hyp[3] << 6;
  auto code_of = [&](uint32_t softsym) -> uint32_t { return (hyp_lut >> (2 * ((softsym >> 16) & 3u))) & 3u; };
  // quad q = symbols E[e_base + 4q .. +4) = bits [8q, 8q+8) of the string; word w = quads 4w .. 4w+3
  for (int qb = 0; qb < 4 * nwords; qb += (int)blockDim.x) {   // (uniform trip count: the shuffles below need every lane)
    const int q = qb + (int)threadIdx.x, w = q >> 2;
    uint32_t piece = 0;
    if (w < nwords) {
      const int64_t e0 = e_base + (int64_t)16 * w;
      if (e0 >= 32 && (uint64_t)(e0 - 32 + 16) <= nsym) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.symbols + (e0 - 32)) + (q & 3));
        piece = (code_of(v.x) << 6 | code_of(v.y) << 4 | code_of(v.z) << 2 | code_of(v.w)) << (24 - 8 * (q & 3));
      } else {
        for (int i = 0; i < 4; ++i) {
          const int64_t e = e0 + 4 * (q & 3) + i;
          uint32_t code = 0;
          if (e < 0) code = 0;                       // (in front of the carried register: never part of a group)
          else if (e < 32) code = (uint32_t)(a.reg_in >> (2 * (31 - e))) & 3u;
          else if ((uint64_t)(e - 32) < nsym) code = code_of(a.symbols[e - 32]);
          piece |= code << (30 - 2 * (4 * (q & 3) + i));
        }
      }
    }
    piece |= __shfl_xor_sync(0xffffffffu, piece, 1);
    piece |= __shfl_xor_sync(0xffffffffu, piece, 2);
    if (w < nwords && (q & 3) == 0) s_bits[w] = piece;
  }
  if (threadIdx.x < 4) s_bits[nwords + threadIdx.x] = 0;   // (a window may look one word past the last register)
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
  const bool half_rate = (pp == 1 && half == 1);
  const uint64_t poly0 = a.deconv[0];
  for (uint32_t t = threadIdx.x; t < nb; t += blockDim.x) {
    const uint64_t j = b0 + t;
    unsigned byte = 0;
    int64_t bit = (int64_t)8 * j;
    if (half_rate && bit >= a.n_out) {
      // Rate 1/2: bit i of the byte is group g+i, register E[K+i .. K+i+32): eight 64-bit strings 2 bits apart,
      // all inside the 128-bit window that starts at the word of the first one.
      const int64_t bo = 2 * (k0 + (bit - a.n_out) - e_base);
      const int wi = (int)(bo >> 5), sh = (int)(bo & 31);
      const uint64_t w0 = ((uint64_t)s_bits[wi] << 32) | s_bits[wi + 1], w1 = ((uint64_t)s_bits[wi + 2] << 32) | s_bits[wi + 3];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int s = sh + 2 * i;   // 0 .. 44
        const uint64_t reg = s ? (w0 << s) | (w1 >> (64 - s)) : w0;
        byte = (byte << 1) | par64(reg & poly0);
      }
      a.out[j] = (uint8_t)byte;
      continue;
    }
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

// ================================================================= MPEG sync
// search_sync (dvb.h:798-840) on the window starting at bytes[pos]; the 204 byte
// offsets are examined by 204 threads, the lowest offset that qualifies wins
// (the reference scans i = 0..203 and stops at the first hit).  Called by the
// whole block; returns the number of bytes to skip (0 = no lock) to every thread.
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

constexpr int kSyncTile = 2048;   // words of the bad-sync mask staged per round (65 536 packets)

__global__ void __launch_bounds__(256)
k_sync_track(const uint8_t *bytes, uint64_t nbytes, const SyncState *st_in,
             const uint32_t *bad_words, uint64_t npackets_flagged, SyncResult *res) {
  __shared__ int s_best, s_pol[2], s_ph[2];
  __shared__ uint32_t s_words[kSyncTile];
  __shared__ int s_stop;
  SyncState st = *st_in;          // every thread keeps an identical copy
  SyncResult r;
  r.consumed = 0; r.produced = 0; r.need_next_sync = 0; r.events = 0;
  auto event = [&](int v, uint64_t pos) {
    if (r.events < 16) { r.event_val[r.events] = v; r.event_pos[r.events] = pos; }
    ++r.events;
  };
  if (st.report_state) { event(0, 0); st.report_state = 0; }
  if (st.synchronized) {
    // run_decoding (dvb.h:842-874): walk the mask until the lock times out.  The block stages the mask a tile at a
    // time; thread 0 walks it in shared memory (round 2: one dependent global load per word, 0.19 ms per batch), and
    // a tile without a single bad sync -- the steady state -- is taken in one step.
    uint64_t p = 0;                 // (thread 0's walk; the others only stage)
    bool unlocked = false;
    const uint64_t nwords = (npackets_flagged + 31) >> 5;
    for (uint64_t w0 = 0; w0 < nwords; w0 += kSyncTile) {
      const uint32_t nw = (uint32_t)min((uint64_t)kSyncTile, nwords - w0);
      uint32_t any = 0;
      for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) { const uint32_t w = bad_words[w0 + i]; s_words[i] = w; any |= w; }
      const int anybad = __syncthreads_or(any != 0);   // (also publishes s_words)
      if (threadIdx.x == 0) {
        const uint64_t tile_end = min(npackets_flagged, (w0 + nw) << 5);   // packets [p, tile_end), p == 32*w0 here
        if (!anybad) {
          const uint64_t nfull = (tile_end - p) >> 5;   // whole words of good packets: 32 good packets each
          if (nfull) { st.lock_timeleft = 3; st.locktime += 32 * nfull; p += 32 * nfull; }
        }
        while (p < tile_end) {
          const uint32_t w = s_words[(p >> 5) - w0];
          const uint64_t lim = min(tile_end, (p & ~(uint64_t)31) + 32);
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
        s_stop = unlocked ? 1 : 0;
      }
      __syncthreads();
      if (s_stop) break;
    }
    if (threadIdx.x != 0) return;
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
// derandomizer::run (dvb.h:1131-1158).  pos_p = 188*((p - r_p) mod 8) with r_p the last packet <= p whose first byte
// is an inverted sync (0xB8 or 0xB8^0x55); before the first reset the carried position keeps cycling.  A packet is kept
// when its de-randomised head is 0x47; kept packets are written back to back.
//
// Three small grids over tiles of 1024 packets (one thread per packet, consecutive threads on consecutive packets):
//   k_derand_tiles  per tile: the last reset, the RS error sum, the packets kept after the tile's first reset, and --
//                   for the packets in front of it, whose pattern position depends on earlier tiles -- the number
//                   kept under each of the 8 possible positions of the tile's first packet;
//   k_derand_chain  one CTA: the position of every tile's first packet (max-scan of the resets), hence its kept
//                   count and its base in the output (sum-scan); the batch totals;
//   k_derand_index  per tile again: output index and pattern position of every packet.
// History: one CTA walking 1024 packets per round took 0.18 ms per 65 536 packets (a load latency and four barriers
// per round); one CTA with a contiguous run of 64 packets per thread 0.16 ms on B200 (a quarter of a million
// uncoalesced accesses through one SM's load/store unit).  Tile records: kDrRec words per tile in a.scratch behind
// the 2 * npackets per-packet words: [0] last reset in the tile + 1 (0: none), [1] kept after the first reset,
// [2..9] kept in front of it by position, [10] RS errors, [11] position of the first packet, [12] output base.
constexpr int kDrTile = 1024;
constexpr int kDrRec = 16;

// inclusive block scans over 1024 threads (s_a / s_b: 32 entries each); every thread must call
__device__ __forceinline__ void block_scan_max_sum(long long &mx, unsigned &sum, long long *s_a, unsigned *s_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, mx, o);
    const unsigned c = __shfl_up_sync(0xffffffffu, sum, o);
    if (lane >= o) { if (v > mx) mx = v; sum += c; }
  }
  __syncthreads();   // (the previous use of s_a / s_b is over)
  if (lane == 31) { s_a[warp] = mx; s_b[warp] = sum; }
  __syncthreads();
  if (warp == 0) {
    long long v = s_a[lane]; unsigned c = s_b[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const long long v2 = __shfl_up_sync(0xffffffffu, v, o);
      const unsigned c2 = __shfl_up_sync(0xffffffffu, c, o);
      if (lane >= o) { if (v2 > v) v = v2; c += c2; }
    }
    s_a[lane] = v; s_b[lane] = c;
  }
  __syncthreads();
  if (warp > 0) { if (s_a[warp - 1] > mx) mx = s_a[warp - 1]; sum += s_b[warp - 1]; }
}

__device__ __forceinline__ uint32_t *derand_rec(const DerandArgs &a, uint64_t tile) { return a.scratch + 2 * a.npackets + kDrRec * tile; }

__global__ void __launch_bounds__(1024)
k_derand_tiles(DerandArgs a) {
  __shared__ long long s_last[32];
  __shared__ unsigned s_cnt[32];
  __shared__ unsigned s_pat[8];
  __shared__ unsigned s_front[8], s_err;
  const int tid = threadIdx.x;
  if (tid < 8) { s_pat[tid] = a.pattern[188 * tid]; s_front[tid] = 0; }
  if (tid == 8) s_err = 0;
  __syncthreads();
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + tid;   // (tile = the CTA: kDrTile threads in the library)
  const bool valid = p < a.npackets;
  const unsigned head = valid ? a.rts[188 * p] : 0u;
  const bool reset = valid && (head == 0xb8u || head == (0xb8u ^ 0x55u));
  unsigned err = (valid && a.flags) ? (unsigned)a.flags[2 * p + 1] : 0u;
  for (int o = 16; o; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
  if ((tid & 31) == 0 && err) atomicAdd(&s_err, err);
  long long lr = reset ? (long long)tid : -1;   // last reset of the tile at or before this packet
  unsigned unused = 0;
  block_scan_max_sum(lr, unused, s_last, s_cnt);
  unsigned kept = 0;
  if (valid && lr >= 0) {
    kept = ((head ^ s_pat[(tid - (int)lr) & 7]) == 0x47u) ? 1u : 0u;
  } else if (valid) {
    for (int f = 0; f < 8; ++f)
      if ((head ^ s_pat[(tid + f) & 7]) == 0x47u) atomicAdd(&s_front[f], 1u);
  }
  long long none = -1;
  block_scan_max_sum(none, kept, s_last, s_cnt);   // (also orders the atomics above before the reads below)
  __syncthreads();
  if (tid == (int)blockDim.x - 1) {
    uint32_t *rec = derand_rec(a, blockIdx.x);
    rec[0] = (uint32_t)(lr + 1);
    rec[1] = kept;
    for (int f = 0; f < 8; ++f) rec[2 + f] = s_front[f];
    rec[10] = s_err;
  }
}

__global__ void __launch_bounds__(1024)
k_derand_chain(DerandArgs a, uint32_t ntiles, uint32_t tile /* packets per tile */) {
  __shared__ long long s_last[32];
  __shared__ unsigned s_cnt[32];
  __shared__ long long carry_last;                 // last reset so far (packet index), -1: none yet
  __shared__ unsigned long long carry_kept, carry_errs;
  const int tid = threadIdx.x;
  if (tid == 0) { carry_last = -1; carry_kept = 0; carry_errs = 0; }
  __syncthreads();
  const long long start_phase = a.pos_in / 188;    // packets since the (virtual) last reset
  for (uint32_t base = 0; base < ntiles; base += blockDim.x) {   // (any whole number of warps; 1024 threads in the library)
    const uint32_t i = base + (uint32_t)tid;
    const bool valid = i < ntiles;
    uint32_t *rec = derand_rec(a, valid ? i : 0);
    const long long first = (long long)i * tile;
    const long long mine = (valid && rec[0]) ? first + (long long)rec[0] - 1 : -1;
    unsigned err = valid ? rec[10] : 0u;
    long long incl = mine;
    block_scan_max_sum(incl, err, s_last, s_cnt);          // err: inclusive sum of the tiles' error counts
    long long prev = __shfl_up_sync(0xffffffffu, incl, 1);   // exclusive: the tiles in front of this one
    if ((tid & 31) == 0) prev = (tid >= 32) ? s_last[(tid >> 5) - 1] : -1;
    if (carry_last > prev) prev = carry_last;
    const int phi = (int)((prev >= 0 ? first - prev : first + start_phase) & 7);
    unsigned kept = valid ? rec[1] + rec[2 + phi] : 0u;
    const unsigned mykept = kept;
    long long none = -1;
    block_scan_max_sum(none, kept, s_last, s_cnt);
    if (valid) { rec[11] = (uint32_t)phi; rec[12] = (uint32_t)(carry_kept + kept - mykept); }
    __syncthreads();   // (every thread has read the carries)
    if (tid == (int)blockDim.x - 1) {
      carry_last = incl > carry_last ? incl : carry_last;
      carry_kept += kept;
      carry_errs += err;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const long long last = carry_last;
    int pos_out;
    if (last >= 0) pos_out = (int)(((long long)a.npackets - last) & 7) * 188;
    else pos_out = (int)(((long long)a.npackets + start_phase) & 7) * 188;
    a.counts[0] = carry_kept;
    a.counts[1] = a.npackets - carry_kept;
    a.counts[2] = (uint64_t)pos_out;
    a.counts[3] = carry_errs;
  }
}

__global__ void __launch_bounds__(1024)
k_derand_index(DerandArgs a) {
  __shared__ long long s_last[32];
  __shared__ unsigned s_cnt[32];
  __shared__ unsigned s_pat[8];
  const int tid = threadIdx.x;
  if (tid < 8) s_pat[tid] = a.pattern[188 * tid];
  __syncthreads();
  const uint32_t *rec = derand_rec(a, blockIdx.x);
  const int phi = (int)rec[11];
  const unsigned base = rec[12];
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + tid;   // (tile = the CTA: kDrTile threads in the library)
  const bool valid = p < a.npackets;
  const unsigned head = valid ? a.rts[188 * p] : 0u;
  const bool reset = valid && (head == 0xb8u || head == (0xb8u ^ 0x55u));
  long long lr = reset ? (long long)tid : -1;
  unsigned unused = 0;
  block_scan_max_sum(lr, unused, s_last, s_cnt);
  const int ph = (lr >= 0 ? tid - (int)lr : tid + phi) & 7;
  const bool keep = valid && (head ^ s_pat[ph]) == 0x47u;
  unsigned incl = keep ? 1u : 0u;
  long long none = -1;
  block_scan_max_sum(none, incl, s_last, s_cnt);
  if (valid) {
    // scratch[p] = output index (0xffffffff when dropped), pattern position in scratch[npackets + p]
    a.scratch[p] = keep ? base + incl - 1u : 0xffffffffu;
    a.scratch[a.npackets + p] = (unsigned)(188 * ph);
  }
}
