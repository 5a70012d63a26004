// vit_v1.cuh -- TEST INFRASTRUCTURE: the device code of k_viterbi.cu as it passed the 144 GPU parity tests of commit
// f37a100 on B200 (the predecessor the rewritten k_vit_dev.cuh is compared with under tests/emu/cuda_emu.h).
// Only change: the dynamic shared memory declaration goes through LDVB_DYN_SMEM (so that g++ can compile it).
constexpr int kVitChunk = 128;

struct VitWarp {
  int32_t *cost;        // [2][64]
  uint64_t *path;       // [2][64]
  int32_t *blk_cost;    // [128] branch cost of the blocks of the chunk being decoded
  uint8_t *blk_cs;      // [128] their coded symbols
  const uint8_t *map;
  int shift, bank;
};

// One chunk of 128 FEC blocks for this warp's decoder (update_sync + viterbi_dec::update,
// dvb.h:1353-1364, viterbi.h:196-263).  Returns the sum of quality over blocks >= discr_delay
// (only computed when need_td: the votes look at it on re-sync chunks only).
// The coded symbol and the branch cost of the 128 blocks do not depend on the decoder state:
// the lanes prepare them in parallel (4 blocks each) before the serial walk, which then touches
// shared memory only.
// R12: rate 1/2 (4 labels, 2 predecessors per state): the trellis rows of the lane's two states
// live in registers (they do not change from block to block), so the serial chain of a block is
// metric loads -> compare/select -> path load -> store.
template <bool R12>
__device__ __forceinline__ int32_t vit_chunk_t(const VitArgs &a, VitWarp &w, const uint8_t *t_pred, const uint8_t *t_us,
                                               const uint8_t *l_pred, const uint8_t *l_us, int nb, uint64_t chunk,
                                               bool write_out, bool need_td, int lane) {
  const int discr_delay = 64 / a.bits_in;   // dvb.h:1369
  const uint64_t path_mask = (1ull << a.path_nbits) - 1;
  const int read_shift = (a.path_depth - 1) * a.path_nbits;
  const int bytes_per_chunk = kVitChunk * a.bits_in / 8;
  int32_t td = 0;
  uint64_t outstream = 0; int nout = 0;
  uint8_t *outp = a.out + chunk * bytes_per_chunk;
  {
    // update_sync (dvb.h:1353-1364): coded symbol and cost of every FEC block of the chunk
    const uint32_t *pin = a.symbols + chunk * (uint64_t)kVitChunk * a.nshifts + w.shift;
#pragma unroll
    for (int q = 0; q < kVitChunk / 32; ++q) {
      const int blk = lane + 32 * q;
      const uint32_t *pb = pin + (size_t)blk * a.nshifts;
      unsigned cs = 0; int32_t bcost = 0;
      for (int i = 0; i < a.nshifts; ++i) {
        const uint32_t sw = __ldg(pb + i);
        cs = ((cs << a.bps) | __ldg(w.map + ((sw >> 16) & 0xffu))) & 0xffu;
        bcost += (int32_t)(int16_t)(sw & 0xffffu);
      }
      w.blk_cs[blk] = (uint8_t)cs;
      w.blk_cost[blk] = bcost;
    }
    __syncwarp();
  }
  uint32_t r_tp[2] = {0, 0}, r_tu[2] = {0, 0};
  int r_lp[2][2] = {{0, 0}, {0, 0}}, r_lu[2][2] = {{0, 0}, {0, 0}};
  if (R12) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s = lane + 32 * h;
      r_tp[h] = *reinterpret_cast<const uint32_t *>(t_pred + s * 4);
      r_tu[h] = *reinterpret_cast<const uint32_t *>(t_us + s * 4);
      r_lp[h][0] = l_pred[s * 2]; r_lp[h][1] = l_pred[s * 2 + 1];
      r_lu[h][0] = l_us[s * 2]; r_lu[h][1] = l_us[s * 2 + 1];
    }
  }
  int bank = w.bank;
  for (int blk = 0; blk < kVitChunk; ++blk) {
    const unsigned cs = w.blk_cs[blk];
    const int32_t bcost = w.blk_cost[blk];
    const int32_t *cc = w.cost + bank * 64;
    const uint64_t *pc = w.path + bank * 64;
    int32_t *cn = w.cost + (bank ^ 1) * 64;
    uint64_t *pn = w.path + (bank ^ 1) * 64;
    int32_t my_m[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s = lane + 32 * h;
      int32_t best_m = 0x7fffffff; int best_pred = 0, best_us = 0;
      if (R12) {
        const int p = (int)((r_tp[h] >> (8 * cs)) & 0xffu);
        const int32_t m0 = cc[r_lp[h][0]], m1 = cc[r_lp[h][1]];
        if (p != 65) { best_m = cc[p] + bcost; best_pred = p; best_us = (int)((r_tu[h] >> (8 * cs)) & 0xffu); }
        if (m0 <= best_m) { best_m = m0; best_pred = r_lp[h][0]; best_us = r_lu[h][0]; }
        if (m1 <= best_m) { best_m = m1; best_pred = r_lp[h][1]; best_us = r_lu[h][1]; }
      } else {
      {
        const int p = t_pred[s * a.ncs + cs];
        if (p != 65) {
          const int32_t m = cc[p] + bcost;
          if (m <= best_m) { best_m = m; best_pred = p; best_us = t_us[s * a.ncs + cs]; }
        }
      }
      if (a.ncs != 1) {
        const uint8_t *lp = l_pred + s * nb, *lu = l_us + s * nb;
        for (int k = 0; k < nb; ++k) {
          const int p = lp[k];
          const int32_t m = cc[p];
          if (m <= best_m) { best_m = m; best_pred = p; best_us = lu[k]; }
        }
      }
      }
      uint64_t np = pc[best_pred];
      if (a.path32) np = (uint64_t)(uint32_t)(((uint32_t)np << a.path_nbits) | (uint32_t)best_us);
      else np = (np << a.path_nbits) | (uint64_t)best_us;
      pn[s] = np; my_m[h] = best_m;
    }
    // best state: minimum, first index wins (viterbi.h:239-243)
    const int32_t bm = __reduce_min_sync(0xffffffffu, min(my_m[0], my_m[1]));
    const unsigned e0 = __ballot_sync(0xffffffffu, my_m[0] == bm);
    const unsigned e1 = __ballot_sync(0xffffffffu, my_m[1] == bm);
    const int bs = e0 ? (__ffs((int)e0) - 1) : (32 + __ffs((int)e1) - 1);
    // normalise (viterbi.h:249)
    cn[lane] = my_m[0] - bm; cn[lane + 32] = my_m[1] - bm;
    bank ^= 1;
    if (need_td) {
      // second best: minimum over all states except the best one (duplicates count)
      const int32_t x0 = (lane == bs) ? 0x7fffffff : my_m[0];
      const int32_t x1 = (lane + 32 == bs) ? 0x7fffffff : my_m[1];
      const int32_t b2 = __reduce_min_sync(0xffffffffu, min(x0, x1));
      if (blk >= discr_delay) td += b2 - bm;
    }
    __syncwarp();
    if (write_out) {
      const unsigned result = (unsigned)((pn[bs] >> read_shift) & path_mask);
      outstream = (outstream << a.bits_in) | result;
      nout += a.bits_in;
      while (nout >= 8) {
        if (lane == 0) *outp = (uint8_t)(outstream >> (nout - 8));
        ++outp; nout -= 8;
      }
    }
  }
  w.bank = bank;
  return td;
}

__device__ __forceinline__ void vit_store(VitDecState *dst, const VitWarp &w, int lane) {
  for (int s = lane; s < 64; s += 32) { dst->cost[s] = w.cost[w.bank * 64 + s]; dst->path[s] = w.path[w.bank * 64 + s]; }
  if (lane == 0) { dst->bank = 0; dst->pad = 0; }   // the bank is re-based to 0 on store
}

// R12 instances are rate 1/2: at most 4 decoders (QPSK: 4, BPSK: 2) = 128 threads, 8 CTAs per SM.
template <bool R12>
__global__ void __launch_bounds__(R12 ? 128 : 512, R12 ? 8 : 2)
k_viterbi(VitArgs a, VitSegArgs sg) {
  LDVB_DYN_SMEM(smem);
  // Layout: trellis pred[64*ncs], us[64*ncs]; rescan lists pred[64*nb], us[64*nb];
  // per warp: cost[2][64] int32, path[2][64] u64.
  const int nb = sg.nb;
  uint8_t *t_pred = smem;
  uint8_t *t_us = t_pred + 64 * a.ncs;
  uint8_t *l_pred = t_us + 64 * a.ncs;
  uint8_t *l_us = l_pred + 64 * nb;
  size_t off = ((size_t)128 * a.ncs + (size_t)128 * nb + 15) & ~(size_t)15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = a.nsyncs;
  int32_t *cost_all = reinterpret_cast<int32_t *>(smem + off);
  off += (size_t)nw * 2 * 64 * 4;
  uint64_t *path_all = reinterpret_cast<uint64_t *>(smem + off);
  off += (size_t)nw * 2 * 64 * 8;
  int32_t *totaldiscr = reinterpret_cast<int32_t *>(smem + off);
  off += (size_t)nw * 4;
  int32_t *blk_cost_all = reinterpret_cast<int32_t *>(smem + off);
  off += (size_t)nw * kVitChunk * 4;
  uint8_t *blk_cs_all = smem + off;
  off += (size_t)nw * kVitChunk;
  off = (off + 15) & ~(size_t)15;
  int *s_ctl = reinterpret_cast<int *>(smem + off);   // [0] current_sync

  const uint32_t g = sg.list ? sg.list[blockIdx.x] : blockIdx.x;
  const bool repair = sg.list != nullptr;
  const bool cold = !repair && g != 0;
  const uint64_t c0 = sg.seg_start[g], c1 = sg.seg_start[g + 1];
  const int P = a.resync_period;

  for (int i = threadIdx.x; i < 64 * a.ncs; i += blockDim.x) { t_pred[i] = a.trellis_pred[i]; t_us[i] = a.trellis_us[i]; }
  __syncthreads();
  // Rescan lists: per state, (pred, us) of the largest label of every distinct predecessor,
  // by increasing label.  Walk the labels downwards, keep first sightings, then reverse.
  for (int s = threadIdx.x; s < 64 && a.ncs != 1; s += blockDim.x) {
    unsigned long long seen = 0;
    int k = nb;
    for (int c = a.ncs - 1; c >= 0; --c) {
      const int p = t_pred[s * a.ncs + c];
      if (p == 65 || ((seen >> p) & 1ull)) continue;
      seen |= 1ull << p;
      if (k > 0) { --k; l_pred[s * nb + k] = (uint8_t)p; l_us[s * nb + k] = t_us[s * a.ncs + c]; }
    }
    // (k == 0 here for a regular code: every state has exactly nb distinct predecessors;
    //  unused slots would repeat the first real entry, which changes nothing)
    for (int j = 0; j < k; ++j) { l_pred[s * nb + j] = l_pred[s * nb + k]; l_us[s * nb + j] = l_us[s * nb + k]; }
  }

  VitWarp w;
  w.cost = cost_all + (size_t)warp * 128;
  w.path = path_all + (size_t)warp * 128;
  w.blk_cost = blk_cost_all + (size_t)warp * kVitChunk;
  w.blk_cs = blk_cs_all + (size_t)warp * kVitChunk;
  w.map = a.maps + (size_t)warp * a.nsymbols;
  w.shift = a.shifts[warp];
  w.bank = 0;
  const VitDecState *src = nullptr;
  if (g == 0) src = a.state + warp;
  else if (repair) src = sg.exit + (size_t)(g - 1) * nw + warp;
  if (src) { for (int s = lane; s < 64; s += 32) { w.cost[s] = src->cost[s]; w.path[s] = src->path[s]; } }
  else { for (int s = lane; s < 64; s += 32) { w.cost[s] = 0; w.path[s] = 0; } }   // viterbi.h:133-145
  if (threadIdx.x == 0) {
    if (g == 0) s_ctl[0] = a.ctl->current_sync;
    else if (repair) s_ctl[0] = sg.ctl_exit[g - 1].current_sync;
    else s_ctl[0] = a.ctl->current_sync;       // speculation: the current decoder does not change
  }
  __syncthreads();

  // One loop over "steps" (a single instance of the block walk in the code):
  //   A  (cold, P > 1)  every decoder on the re-sync chunks in front of c0, with their votes.  A decoder
  //      of a WRONG hypothesis is fed noise: its 64 survivors coalesce like a random genealogy (time scale
  //      ~64 blocks, exponential tail), so it needs ~2000 blocks where the right one needs a few dozen.
  //      A segment that starts fewer than warm_others re-sync chunks into the batch does better: the other
  //      decoders only ever run on re-sync chunks, so their state at c0 follows EXACTLY from the carried
  //      state and the (few) re-sync chunks in front of c0.
  //   B  (cold)  the decoder that is current after A restarts cold on the last warm_chunks chunks
  //      (P == 1: every decoder, with the votes).  Nothing is written during A and B; votes ARE taken, from
  //      cold decoders: which hypothesis is current at c0 is the outcome of the last vote before c0.
  //   M  the segment's own chunks.
  uint64_t nA = 0, nB = 0;
  if (cold && (sg.warm_chunks || sg.warm_others)) {
    nB = sg.warm_chunks;
    if (P > 1) {
      const uint64_t first_resync = (uint64_t)((P - sg.phase0 % P) % P);
      nA = (c0 - first_resync) / (uint64_t)P;                     // re-sync chunks in [0, c0)
      if (nA < sg.warm_others) {
        const VitDecState *cs0 = a.state + warp;
        for (int s = lane; s < 64; s += 32) { w.cost[s] = cs0->cost[s]; w.path[s] = cs0->path[s]; }
        w.bank = 0;
      } else {
        nA = sg.warm_others;
      }
    }
  }
  __syncthreads();
  const uint64_t nsteps = nA + nB + (c1 - c0);
  for (uint64_t it = 0; it < nsteps; ++it) {
    uint64_t chunk; bool runs, vote, out, need_td;
    const int current = s_ctl[0];
    if (it < nA) {
      chunk = c0 - (nA - it) * (uint64_t)P; runs = true; vote = true; out = false; need_td = true;
    } else if (it < nA + nB) {
      if (it == nA && P > 1 && warp == current) {
        for (int s = lane; s < 64; s += 32) { w.cost[s] = 0; w.path[s] = 0; }
        w.bank = 0;
        __syncwarp();
      }
      chunk = c0 - (nA + nB - it); runs = (P == 1) || warp == current; vote = (P == 1); out = false; need_td = vote;
    } else {
      if (it == nA + nB) {
        vit_store(sg.entry + (size_t)g * nw + warp, w, lane);
        if (threadIdx.x == 0) {
          VitCtl ce; ce.current_sync = current; ce.resync_phase = (int)(((uint64_t)sg.phase0 + c0) % (uint64_t)P);
          sg.ctl_entry[g] = ce;
        }
      }
      chunk = c0 + (it - nA - nB);
      const bool resync = (((uint64_t)sg.phase0 + chunk) % (uint64_t)P) == 0;
      const bool mine = (warp == current);
      runs = mine || resync; vote = resync; out = mine; need_td = resync;
    }
    if (runs) {
      const int32_t td = vit_chunk_t<R12>(a, w, t_pred, t_us, l_pred, l_us, nb, chunk, out, need_td, lane);
      if (lane == 0) totaldiscr[warp] = td;
    }
    __syncthreads();
    if (threadIdx.x == 0 && vote) {   // dvb.h:1402-1411
      int best = current;
      for (int s = 0; s < a.nsyncs; ++s) if (totaldiscr[s] > totaldiscr[best]) best = s;
      s_ctl[0] = best;
    }
    __syncthreads();
  }
  if (c1 == c0) {   // (cannot happen: every segment owns at least one chunk)
    vit_store(sg.entry + (size_t)g * nw + warp, w, lane);
  }
  vit_store(sg.exit + (size_t)g * nw + warp, w, lane);
  if (threadIdx.x == 0) {
    VitCtl ce; ce.current_sync = s_ctl[0]; ce.resync_phase = (int)(((uint64_t)sg.phase0 + c1) % (uint64_t)P);
    sg.ctl_exit[g] = ce;
  }
}

// entry(g) == exit(g-1), bit for bit, for every decoder; ok[g] and the number of failures.
__global__ void __launch_bounds__(64)
k_vit_verify(VitSegArgs sg, int nsyncs, uint8_t *ok, uint32_t *nfail) {
  const uint32_t g = blockIdx.x + 1;
  if (g >= sg.nseg) return;
  const int s = threadIdx.x;
  bool same = true;
  for (int d = 0; d < nsyncs; ++d) {
    const VitDecState &e = sg.entry[(size_t)g * nsyncs + d];
    const VitDecState &x = sg.exit[(size_t)(g - 1) * nsyncs + d];
    same = same && e.cost[s] == x.cost[s] && e.path[s] == x.path[s];
  }
  if (s == 0)
    same = same && sg.ctl_entry[g].current_sync == sg.ctl_exit[g - 1].current_sync &&
           sg.ctl_entry[g].resync_phase == sg.ctl_exit[g - 1].resync_phase;
  const int all = __syncthreads_and(same ? 1 : 0);
  if (s == 0) {
    ok[g] = (uint8_t)all;
    if (!all) atomicAdd(nfail, 1u);
  }
}

// The last segment's exit state becomes the carried state.
__global__ void k_vit_commit(VitArgs a, VitSegArgs sg) {
  const VitDecState *src = sg.exit + (size_t)(sg.nseg - 1) * a.nsyncs;
  for (int i = threadIdx.x; i < a.nsyncs * 64; i += blockDim.x) {
    a.state[i / 64].cost[i % 64] = src[i / 64].cost[i % 64];
    a.state[i / 64].path[i % 64] = src[i / 64].path[i % 64];
  }
  if (threadIdx.x < a.nsyncs) { a.state[threadIdx.x].bank = 0; a.state[threadIdx.x].pad = 0; }
  if (threadIdx.x == 0) *a.ctl = sg.ctl_exit[sg.nseg - 1];
}
