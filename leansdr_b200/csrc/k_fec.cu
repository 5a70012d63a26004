// k_fec.cu -- K4/K6/K7: algebraic deconvolution, MPEG sync tracking, byte
// re-alignment, de-interleaving, Reed-Solomon RS(204,188) and de-randomisation.
//
// All of this is integer/bit work on a stream that is ~80x smaller than the IQ
// input (0.104 byte per sample at QPSK 1/2), so the kernels are written for
// exactness and full parallelism over bytes/packets, not for HBM peak.
#include "common.cuh"
#include "kernels.h"

namespace ldvb {

namespace {

__device__ __forceinline__ unsigned par64(uint64_t v) { return __popcll(v) & 1; }

// =============================================================== deconvolution
// deconvol_sync::readbyte (dvb.h:369-389) as a position-indexed computation.
// Bit group g (punctperiod bits) is the parity of the 64-bit IQ shift register
// after K_g = k0 + g*(punctweight/2) symbols, k0 = symbols needed to fill the
// carried register to 64 bits.  The register holds the last 32 symbols, two IQ
// bits each, newest in the LSBs.  Output byte j = stream bits [8j, 8j+8) where
// the carried accumulator supplies the first n_out bits.
__device__ __forceinline__ uint64_t deconv_reg(const DeconvArgs &a, int64_t K) {
  // Register after K symbols of this batch: (carry << 2K) | new IQ bits.
  uint64_t reg = (K >= 32) ? 0ull : (a.reg_in << (2 * K));
  const int64_t first = (K > 32) ? K - 32 : 0;
  for (int64_t s = first; s < K; ++s) {
    const uint32_t sym = (a.symbols[s] >> 16) & 3u;
    reg |= (uint64_t)a.hyp[sym] << (2 * (K - 1 - s));
  }
  return reg;
}

#include "k_ctl_fec.cuh"   // k_deconv_tiled, sync_search_window, k_sync_track, k_derand_tiles / _chain / _index

// Carry for the next batch (register, leftover bits, symbols consumed): one thread.
__global__ void k_deconv(DeconvArgs a, uint64_t *carry_out) {
  const uint64_t j = a.nbytes;
  const int pp = a.punctperiod, half = a.punctweight / 2;
  const int64_t k0 = (a.n_in >= 64) ? 0 : (64 - a.n_in) / 2;  // n_in is always even
  if (threadIdx.x == 0 && blockIdx.x == 0 && carry_out) {
    (void)j;
    // Groups completed: the last byte ends at stream bit 8*nbytes-1.
    const int64_t total_bits = (int64_t)8 * a.nbytes;
    int64_t ngroups = 0;
    if (total_bits > a.n_out) ngroups = (total_bits - a.n_out + pp - 1) / pp;
    uint64_t reg = a.reg_in, acc = a.out_acc;
    int n_in = a.n_in, n_out = a.n_out;
    int64_t consumed = 0;
    if (ngroups > 0) {
      consumed = k0 + (ngroups - 1) * half;
      reg = deconv_reg(a, consumed);
      n_in = 64 - a.punctweight;
      // Leftover bits: low bits of the last group.
      const int64_t produced_bits = a.n_out + ngroups * pp;
      n_out = (int)(produced_bits - total_bits);
      acc = 0;
      for (int b = pp - 1; b >= 0; --b) acc = (acc << 1) | par64(reg & a.deconv[b]);
    } else {
      n_out = (int)(a.n_out - total_bits);
    }
    carry_out[0] = reg; carry_out[1] = (uint64_t)(int64_t)n_in;
    carry_out[2] = acc; carry_out[3] = (uint64_t)(int64_t)n_out;
    carry_out[4] = (uint64_t)consumed;
  }
}

// ============================================== --hs: dvb_deconvol_sync_hard
// dvb_deconvol_sync<u8>::run (dvb.h:633-660) over deconvol_poly2<u8,uint32_t,uint64_t,0x3ba,0x38f70>
// (convolutional.h:96-187).  The reference convolves 32 symbols at a time with bit-sliced shift
// registers; written out per position it is a GF(2) FIR over the remapped I/Q bits:
//     decoded bit m = XOR_b  PD_I[b] & I[m-b]  ^  PD_Q[b] & Q[m-b]        (taps b = 0..4,  poly 0x3ba)
//     error   bit m = XOR_b  PE_I[b] & I[m-b]  ^  PE_Q[b] & Q[m-b]        (taps b = 2..8,  poly 0x38f70)
// with poly bit 2b+1 = I tap b, bit 2b = Q tap b, and (I,Q) = lut[alignment][symbol].  Errors are
// counted over the second half of a 64-byte chunk (bits 256..511) for all four alignments on every
// resync_period-th chunk; the best one decodes from the NEXT chunk on.  Everything is a function of
// the position, so the chunks are processed in parallel; only the "which alignment is locked" chain
// is walked by one thread over the (few) resync chunks.
__device__ __constant__ uint8_t kHsLut[4][4] = {{0, 1, 2, 3}, {2, 0, 3, 1}, {1, 0, 3, 2}, {0, 2, 1, 3}};   // dvb.h:674-705

// Remapped (I,Q) bit pair of the symbol at stream position n (n < 0: the carried history).
__device__ __forceinline__ uint32_t hs_iq(const HsDeconvArgs &a, int64_t n, int sync) {
  uint32_t sym;
  if (n >= 0) sym = (__ldg(a.symbols + n) >> 16) & 3u;
  else {
    const int back = (int)(-n);                    // 1 = the symbol right in front of symbols[0]
    if (back > a.hist_valid || back > 32) return 0; // registers start at zero (convolutional.h:92)
    sym = (uint32_t)(a.hist >> (2 * (back - 1))) & 3u;
  }
  return kHsLut[sync][sym];
}

__device__ __forceinline__ uint32_t hs_fir_bit(const HsDeconvArgs &a, int64_t m, int sync, uint32_t poly, int ntaps) {
  uint32_t acc = 0;
  for (int b = 0; b < ntaps; ++b) {
    const uint32_t sel = (poly >> (2 * b)) & 3u;   // bit 1: I tap, bit 0: Q tap
    if (sel) acc ^= hs_iq(a, m - b, sync) & sel;
  }
  return (__popc(acc) & 1u);
}

// One CTA (4 warps = 4 alignments) per resync chunk: errors over bits 256..511 of the chunk.
__global__ void __launch_bounds__(128) k_hs_errors(HsDeconvArgs a, uint64_t first_resync, uint32_t ngroups) {
  const uint32_t g = blockIdx.x;
  if (g >= ngroups) return;
  const int sync = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c = (int64_t)first_resync + (int64_t)g * a.resync_period;
  uint32_t err = 0;
  for (int k = 0; k < 8; ++k) err += hs_fir_bit(a, c * 512 + 256 + lane * 8 + k, sync, 0x38f70u, 9);
  for (int o = 16; o; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
  if (lane == 0) a.errors[g * 4 + sync] = err;
}

// The lock chain (dvb.h:640-657), walked over the resync chunks only: lock_of_chunk[g] = alignment
// in force AFTER the vote of group g (the chunks behind the voting chunk, up to the next vote).
__global__ void k_hs_lock(HsDeconvArgs a, uint32_t ngroups) {
  if (threadIdx.x || blockIdx.x) return;
  int locked = a.locked;
  for (uint32_t g = 0; g < ngroups; ++g) {
    const uint32_t *e = a.errors + (size_t)g * 4;
    int best = 0; uint32_t eb = e[0];
    for (int s = 1; s < 4; ++s) if (e[s] < eb) { eb = e[s]; best = s; }
    locked = best;
    a.lock_of_chunk[g] = (uint8_t)locked;
  }
  a.state_out[0] = locked;
}

// Thread per output byte: the 12 symbols 8j-4 .. 8j+7 are remapped once into two bit strings, the
// eight decoded bits are parities of shifted windows (taps b = 0..4 of poly 0x3ba).
__global__ void __launch_bounds__(256) k_hs_decode(HsDeconvArgs a, uint64_t first_resync) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.nchunks * 64) return;
  // chunk c is decoded with the alignment locked BEFORE its own vote
  const uint64_t c = j >> 6;
  int sync = a.locked;
  if (c >= first_resync) {
    const uint64_t g = (c - first_resync) / (uint64_t)a.resync_period;
    const bool voting = (c - first_resync) % (uint64_t)a.resync_period == 0;
    if (!voting) sync = a.lock_of_chunk[g];
    else if (g) sync = a.lock_of_chunk[g - 1];
  }
  uint32_t I = 0, Q = 0;      // bit i = symbol 8j - 4 + i
  if (j) {
    const uint4 *w = reinterpret_cast<const uint4 *>(a.symbols + 8 * j - 4);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const uint4 v = __ldg(w + q);
      const uint32_t sy[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t iq = kHsLut[sync][(sy[t] >> 16) & 3u];
        I |= (iq >> 1) << (4 * q + t); Q |= (iq & 1u) << (4 * q + t);
      }
    }
  } else {
    for (int i = 0; i < 12; ++i) {
      const uint32_t iq = hs_iq(a, (int64_t)i - 4, sync);
      I |= (iq >> 1) << i; Q |= (iq & 1u) << i;
    }
  }
  // decoded bit k (position 8j + k) = XOR_b PD_I[b] & I[k + 4 - b] ^ PD_Q[b] & Q[k + 4 - b]
  uint32_t byte = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t acc = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      const uint32_t sel = (0x3bau >> (2 * b)) & 3u;
      if (sel & 2u) acc ^= (I >> (k + 4 - b)) & 1u;
      if (sel & 1u) acc ^= (Q >> (k + 4 - b)) & 1u;
    }
    byte = (byte << 1) | acc;
  }
  a.out[j] = (uint8_t)byte;
}

// ================================================================= MPEG sync
// mpeg_sync (dvb.h:742-874).  Locked tracking is data parallel (k_sync_flags +
// a word-wise scan of the bad-sync mask); acquisition (run_searching,
// search_sync) is a sequential window walk done by one thread.

__device__ __forceinline__ unsigned realigned(const uint8_t *b, uint64_t i, int bitphase, int polarity) {
  const unsigned w = ((unsigned)b[i] << 8) | b[i + 1];
  return ((w >> bitphase) ^ (unsigned)polarity) & 0xffu;
}

__global__ void __launch_bounds__(256)
k_sync_flags(const uint8_t *bytes, uint64_t npackets, const SyncState *st, uint32_t *bad_words) {
  // One thread per packet; 32 packets per mask word (ballot).
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  if (p < npackets) {
    const int phase8 = (int)((st->phase8 + p) & 7);
    const unsigned expected = phase8 ? 0x47u : 0xb8u;
    bad = realigned(bytes, 204 * p, st->bitphase, st->polarity) != expected;
  }
  const unsigned m = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && (p >> 5) <= ((npackets + 31) >> 5)) bad_words[p >> 5] = m;
}

__global__ void __launch_bounds__(256)
k_realign(const uint8_t *bytes, uint64_t n, int bitphase, int polarity, uint8_t *out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint8_t)realigned(bytes, i, bitphase, polarity);
}

// ============================================== de-interleaver + Reed-Solomon
struct Gf {
  const uint8_t *ex, *lg;
  __device__ __forceinline__ unsigned mul(unsigned x, unsigned y) const {
    if (!x || !y) return 0;
    return ex[lg[x] + lg[y]];
  }
  __device__ __forceinline__ unsigned div(unsigned x, unsigned y) const {  // rs.h:68-72
    if (!x) return 0;
    return ex[lg[x] + 255 - lg[y]];
  }
  __device__ __forceinline__ unsigned inv(unsigned x) const { return ex[255 - lg[x]]; }
};

// Syndromes S_j = P(alpha^j), j = 0..15, P = sum_i r[i] x^(203-i) (rs.h:116-130).
// Each lane folds its bytes (i = lane + 32k), lanes are combined by XOR.
__device__ __forceinline__ bool rs_syndromes(const Gf &gf, const uint8_t *r, int lane, uint8_t synd[16]) {
  unsigned acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0;
  for (int i = lane; i < 204; i += 32) {
    const unsigned v = r[i];
    if (v) {
      const unsigned lv = gf.lg[v];
      const unsigned e = 203 - i;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] ^= gf.ex[(lv + j * e) % 255];
    }
  }
  bool any = false;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    unsigned v = acc[j];
    v ^= __shfl_xor_sync(0xffffffffu, v, 16);
    v ^= __shfl_xor_sync(0xffffffffu, v, 8);
    v ^= __shfl_xor_sync(0xffffffffu, v, 4);
    v ^= __shfl_xor_sync(0xffffffffu, v, 2);
    v ^= __shfl_xor_sync(0xffffffffu, v, 1);
    synd[j] = (uint8_t)v;
    any |= (v != 0);
  }
  return any;
}

__device__ __forceinline__ unsigned poly_eval(const Gf &gf, const uint8_t *poly, int deg, unsigned x) {
  unsigned acc = 0;  // rs.h:133-138
  for (; deg >= 0; --deg) acc = gf.mul(acc, x) ^ poly[deg];
  return acc;
}

// rs_engine::correct (rs.h:176-268) for one packet held in shared memory.
// Berlekamp-Massey and Omega run redundantly on every lane (no divergence); the
// root scan alpha^0..alpha^254 is split across lanes (the error locator has at
// most L roots, so scanning all candidates equals the reference's early exit).
__device__ bool rs_correct(const Gf &gf, uint8_t *pin, uint8_t *pout, uint8_t synd[16], int lane,
                           int *bits_corrected) {
  uint8_t C[17], B[17], T[16];
#pragma unroll
  for (int i = 0; i < 17; ++i) { C[i] = 0; B[i] = 0; }
  C[0] = 1; B[0] = 1;
  int L = 0, m = 1;
  unsigned b = 1;
  for (int n = 0; n < 16; ++n) {
    unsigned d = synd[n];
    for (int i = 1; i <= L; ++i) d ^= gf.mul(C[i], synd[n - i]);
    if (!d) {
      ++m;
    } else if (2 * L <= n) {
      for (int i = 0; i < 16; ++i) T[i] = C[i];
      const unsigned k = gf.inv(b);
      for (int i = 0; i < 16 - m; ++i) C[m + i] ^= gf.mul(d, gf.mul(k, B[i]));
      L = n + 1 - L;
      for (int i = 0; i < 16; ++i) B[i] = T[i];
      b = d;
      m = 1;
    } else {
      const unsigned k = gf.inv(b);
      for (int i = 0; i < 16 - m; ++i) C[m + i] ^= gf.mul(d, gf.mul(k, B[i]));
      ++m;
    }
  }
  uint8_t omega[16], Cp[15];
  for (int i = 0; i < 16; ++i) omega[i] = 0;
  for (int i = 0; i < 16; ++i)
    for (int j = 0; i + j < 16; ++j) omega[i + j] ^= gf.mul(synd[i], C[j]);
  for (int i = 0; i < 15; ++i) Cp[i] = (i & 1) ? 0 : C[i + 1];
  int nbits = 0;
  for (int i = lane; i < 255; i += 32) {
    const unsigned r = gf.ex[i];
    if (poly_eval(gf, C, L, r) == 0) {
      const unsigned xk = gf.inv(r);
      const int loc = (255 - i) % 255;
      if (loc < 204) {
        const unsigned num = gf.mul(xk, poly_eval(gf, omega, L < 16 ? L : 15, r));
        const unsigned den = poly_eval(gf, Cp, 14, r);
        const unsigned e = gf.div(num, den);
        nbits += __popc(e);
        if (loc >= 16) pout[203 - loc] ^= (uint8_t)e;
        pin[203 - loc] ^= (uint8_t)e;
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) nbits += __shfl_xor_sync(0xffffffffu, nbits, o);
  *bits_corrected += nbits;
  __syncwarp();
  return rs_syndromes(gf, pin, lane, synd);
}

template <bool DEINT>
__global__ void __launch_bounds__(128)
k_rs(const uint8_t *src, uint64_t npackets, const uint8_t *gexp, const uint8_t *glog,
     uint8_t *rs_out, uint8_t *rts_out, int32_t *flags) {
  __shared__ uint8_t s_exp[512], s_log[256];
  __shared__ uint8_t s_pkt[4][208], s_msg[4][192];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s_exp[i] = gexp[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_log[i] = glog[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Gf gf{s_exp, s_log};
  for (uint64_t p = (uint64_t)blockIdx.x * 4 + warp; p < npackets; p += (uint64_t)gridDim.x * 4) {
    uint8_t *pin = s_pkt[warp], *pout = s_msg[warp];
    for (int i = lane; i < 204; i += 32) {
      uint8_t v;
      if (DEINT) v = src[204 * (p + (uint64_t)(i % 12)) + i];  // dvb.h:936-941: in[2244+204p+i-12*delay_i]
      else v = src[204 * p + i];
      pin[i] = v;
      if (i < 188) pout[i] = v;  // dvb.h:1011-1014
    }
    __syncwarp();
    if (DEINT && rs_out)
      for (int i = lane; i < 204; i += 32) rs_out[204 * p + i] = pin[i];
    uint8_t synd[16];
    bool corrupted = rs_syndromes(gf, pin, lane, synd);
    int nerr = 0;
    if (corrupted) corrupted = rs_correct(gf, pin, pout, synd, lane, &nerr);
    __syncwarp();
    if (corrupted && lane == 0) pout[0] ^= 0x55;  // dvb.h:1045
    __syncwarp();
    for (int i = lane; i < 188; i += 32) rts_out[188 * p + i] = pout[i];
    if (lane == 0 && flags) { flags[2 * p] = corrupted ? 1 : 0; flags[2 * p + 1] = nerr; }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
k_derand_out(DerandArgs a) {
  // 64 threads (one per 3 bytes... simply byte-wise) per packet, 4 packets per block.
  const uint64_t p = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 6);
  if (p >= a.npackets) return;
  const unsigned idx = a.scratch[p];
  if (idx == 0xffffffffu || idx >= a.ts_cap) return;
  const unsigned pos = a.scratch[a.npackets + p];
  const uint8_t *src = a.rts + 188 * p;
  uint8_t *dst = a.ts_out + 188 * (uint64_t)idx;
  for (int i = threadIdx.x & 63; i < 188; i += 64) dst[i] = src[i] ^ a.pattern[pos + i];
}

}  // namespace

cudaError_t launch_hs_deconv(const HsDeconvArgs &a, cudaStream_t st, int *launches) {
  *launches = 0;
  if (!a.nchunks) return cudaSuccess;
  // first chunk of the batch on which all alignments vote (resync_phase == 0)
  const uint64_t first = (uint64_t)((a.resync_period - a.resync_phase) % a.resync_period);
  const uint32_t ngroups = first < a.nchunks ? (uint32_t)((a.nchunks - first + a.resync_period - 1) / a.resync_period) : 0;
  if (ngroups) { k_hs_errors<<<ngroups, 128, 0, st>>>(a, first, ngroups); ++*launches; }
  k_hs_lock<<<1, 32, 0, st>>>(a, ngroups); ++*launches;
  k_hs_decode<<<(unsigned)((a.nchunks * 64 + 255) / 256), 256, 0, st>>>(a, first); ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_deconv_carry(const DeconvArgs &a, uint64_t nsym, uint64_t *carry_out, cudaStream_t st) {
  if (a.nbytes) k_deconv_tiled<<<(unsigned)((a.nbytes + kDcBytes - 1) / kDcBytes), 256, 0, st>>>(a, nsym);
  k_deconv<<<1, 32, 0, st>>>(a, carry_out);
  return cudaGetLastError();
}

cudaError_t launch_sync_flags(const uint8_t *bytes, uint64_t npackets, const SyncState *st_dev,
                              uint32_t *bad_words, cudaStream_t st) {
  if (!npackets) return cudaSuccess;
  k_sync_flags<<<(unsigned)((npackets + 255) / 256), 256, 0, st>>>(bytes, npackets, st_dev, bad_words);
  return cudaGetLastError();
}

cudaError_t launch_sync_track(const uint8_t *bytes, uint64_t nbytes, const SyncState *st_dev,
                              const uint32_t *bad_words, uint64_t npackets_flagged, SyncResult *res,
                              cudaStream_t st) {
  k_sync_track<<<1, 256, 0, st>>>(bytes, nbytes, st_dev, bad_words, npackets_flagged, res);
  return cudaGetLastError();
}

cudaError_t launch_realign(const uint8_t *bytes, uint64_t n, int bitphase, int polarity, uint8_t *out,
                           cudaStream_t st) {
  if (!n) return cudaSuccess;
  k_realign<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bytes, n, bitphase, polarity, out);
  return cudaGetLastError();
}

cudaError_t launch_deint_rs(const DeintRsArgs &a, cudaStream_t st) {
  if (!a.npackets) return cudaSuccess;
  unsigned blocks = (unsigned)((a.npackets + 3) / 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_rs<true><<<blocks, 128, 0, st>>>(a.mpeg, a.npackets, a.gf_exp, a.gf_log, a.rs_out, a.rts_out, a.flags);
  return cudaGetLastError();
}

cudaError_t launch_rs_only(const RsOnlyArgs &a, cudaStream_t st) {
  if (!a.npackets) return cudaSuccess;
  unsigned blocks = (unsigned)((a.npackets + 3) / 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_rs<false><<<blocks, 128, 0, st>>>(a.rs_in, a.npackets, a.gf_exp, a.gf_log, nullptr, a.rts_out, a.flags);
  return cudaGetLastError();
}

cudaError_t launch_derand(const DerandArgs &a, cudaStream_t st, int *launches) {
  if (!a.npackets) {
    // counts must still be defined
    uint64_t z[4] = {0, 0, (uint64_t)a.pos_in, 0};
    return cudaMemcpyAsync(a.counts, z, sizeof(z), cudaMemcpyHostToDevice, st);
  }
  const uint32_t ntiles = (uint32_t)((a.npackets + kDrTile - 1) / kDrTile);
  k_derand_tiles<<<ntiles, kDrTile, 0, st>>>(a);
  k_derand_chain<<<1, 1024, 0, st>>>(a, ntiles, kDrTile);
  k_derand_index<<<ntiles, kDrTile, 0, st>>>(a);
  k_derand_out<<<(unsigned)((a.npackets + 3) / 4), 256, 0, st>>>(a);
  if (launches) *launches += 4;
  return cudaGetLastError();
}

}  // namespace ldvb
