// emu_vit.cpp -- TEST INFRASTRUCTURE.  Runs k_viterbi of leansdr_b200/csrc/k_vit_dev.cuh on the host (cuda_emu.h) and
// compares output bytes, entry / exit states and hypothesis control, bit for bit, with its predecessor (vit_v1.cuh:
// the text that ran on B200 under the GPU parity tests of commit f37a100).  The stream is made here from the product's
// own host trellis (tables.cpp): a random walk through the code for hypothesis 0 with symbol errors and random
// (negative) costs -- signal for one decoder, noise for the others -- so that labelled branches, rescan winners,
// unique best states, ties and the all-tied cold start all occur; with `noise` every decoder sees noise.
// With a sixth argument "oracle" the checker is the ORACLE instead (oracle/dvbs_oracle.c, pinned to the reference's
// viterbi_sync): every hypothesis of the configuration, one serial segment from the constructor state; output bytes,
// the metrics and path registers of every decoder at the end and the elected hypothesis must equal orc_viterbi_run's.
// Usage: emu_vit <fec 0..5> <mode: full|generic|r12|ws> <seed> [noise|signal] [layout 0..2] [oracle]; exit code 0 = identical.
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

#include "../../leansdr_b200/csrc/kernels.h"
#include "../../leansdr_b200/csrc/tables.h"
#include "../../include/leandvb_b200.h"
extern "C" {
#include "../../oracle/dvbs_oracle.h"
}

namespace ldvb {
namespace v1 {
#include "vit_v1.cuh"
}
namespace v2 {
#include "../../leansdr_b200/csrc/k_vit_dev.cuh"
}
int vit_rescan_entries(int bits_in) { return bits_in >= 6 ? 64 : (1 << bits_in); }
}  // namespace ldvb

using namespace ldvb;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { fprintf(stderr, "MISMATCH %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++g_fail; } } while (0)

static size_t smem_bytes(int ncs, int nb, int nsyncs) {
  return (((size_t)128 * ncs + (size_t)128 * nb + 15) & ~(size_t)15) + (size_t)nsyncs * (2 * 64 * 4 + 2 * 64 * 8 + 4 + 128 * 5) + 96;
}

struct Run {
  std::vector<uint8_t> out;
  std::vector<VitDecState> entry, exit, state;
  std::vector<VitCtl> ctl_entry, ctl_exit;
  VitCtl ctl;
};

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: emu_vit <fec> <full|generic|r12> <seed> [noise]\n"); return 2; }
  const int fec = atoi(argv[1]);
  const std::string mode = argv[2];
  const uint64_t seed = strtoull(argv[3], nullptr, 10);
  const bool noise_only = argc > 4 && std::string(argv[4]) == "noise";
  // layout 0: two cold segments far from the start (phases A and B); 1: the first cold segment starts fewer than
  // warm_others re-sync chunks into the batch (other decoders exactly from the carried state); 2: resync_period 1
  // (--fastlock: every decoder on every chunk, a vote per chunk)
  const int layout = argc > 5 ? atoi(argv[5]) : 0;
  const bool vs_oracle = argc > 6 && std::string(argv[6]) == "oracle";
  const int P = layout == 2 ? 1 : 8;
  std::mt19937_64 rng(seed);

  Trellis tr;
  if (!make_trellis(fec, &tr)) { fprintf(stderr, "no trellis\n"); return 2; }
  const Cstln cst = make_cstln(LDVB_CSTLN_QPSK, fec, false);
  const VitSyncs vs = make_vitsyncs(cst, tr);
  const int nsyncs = vs_oracle ? vs.nsyncs : std::min(vs.nsyncs, mode == "ws" ? 4 : 3);   // (one OS thread per CUDA thread: a few decoders are enough)
  const int nsh = vs.nshifts, bps = vs.bps, ncs = tr.ncs;
  const int nb = vit_rescan_entries(tr.bits_in);
  bool full = nb == 64;
  for (int s = 0; s < 64 && full; ++s) {
    unsigned long long seen = 0;
    for (int c = 0; c < ncs; ++c) { const int p = tr.pred[s * ncs + c]; if (p < 64) seen |= 1ull << p; }
    full = seen == ~0ull;
  }
  if (mode == "full" && !full) { fprintf(stderr, "trellis of fec %d is not full\n", fec); return 2; }

  // forward branches of the code: from state p, (next state, label)
  std::vector<std::vector<std::pair<int, int>>> fwd(64);
  for (int s = 0; s < 64; ++s) for (int c = 0; c < ncs; ++c) { const int p = tr.pred[s * ncs + c]; if (p < 64) fwd[p].push_back({s, c}); }
  std::vector<int> inv(1 << bps, 0);   // coded bits -> symbol index under hypothesis 0
  for (int sym = 0; sym < cst.nsymbols; ++sym) inv[vs.map[0][sym] & ((1 << bps) - 1)] = sym;

  const uint64_t nchunks = 32;
  const uint64_t nblocks = nchunks * 128;
  std::vector<uint32_t> symbols(nblocks * nsh + 64, 0);
  {
    int state = 0;
    const int sh0 = vs.shift[0];
    for (uint64_t b = 0; b < nblocks; ++b) {
      const auto &br = fwd[state][rng() % fwd[state].size()];
      state = br.first;
      for (int i = 0; i < nsh; ++i) {
        const int bits = (br.second >> (bps * (nsh - 1 - i))) & ((1 << bps) - 1);
        int sym = inv[bits];
        if (noise_only || rng() % 40 == 0) sym = (int)(rng() % cst.nsymbols);
        const int cost = -(int)(rng() % 400) - ((rng() % 16 == 0) ? 0 : 1);   // negative, now and then 0
        symbols[b * nsh + sh0 + i] = (uint32_t)sym << 16 | ((uint32_t)cost & 0xffffu);
      }
    }
    for (auto &w : symbols) if (!w) w = (uint32_t)(rng() % cst.nsymbols) << 16 | ((uint32_t)(-(int)(rng() % 400)) & 0xffffu);
  }
  std::vector<uint8_t> maps;
  for (int d = 0; d < nsyncs; ++d) maps.insert(maps.end(), vs.map[d].begin(), vs.map[d].end());
  std::vector<int32_t> shifts(vs.shift.begin(), vs.shift.begin() + nsyncs);

  const std::vector<uint64_t> seg_start = layout == 1 ? std::vector<uint64_t>{0, 8, 24, nchunks}
                                        : layout == 2 ? std::vector<uint64_t>{0, 9, 20, nchunks}
                                                      : std::vector<uint64_t>{0, 16, 24, nchunks};   // P = 8: boundaries on re-sync chunks
  const uint32_t nseg = (uint32_t)seg_start.size() - 1;
  const size_t smem = smem_bytes(ncs, nb, nsyncs);

  auto run = [&](int version, Run &r) {
    r.out.assign(nchunks * 16 * tr.bits_in, 0xee);
    r.entry.assign((size_t)nseg * nsyncs, VitDecState{}); r.exit = r.entry;
    r.state.assign(nsyncs, VitDecState{});
    // a carried state that is not all zero: decoder d starts from a few metrics and paths
    std::mt19937_64 r2(seed * 77 + 5);
    for (auto &st : r.state) for (int s = 0; s < 64; ++s) { st.cost[s] = (int32_t)(r2() % 300); st.path[s] = r2(); }
    for (auto &st : r.state) st.cost[(int)(r2() % 64)] = 0;
    r.ctl_entry.assign(nseg, VitCtl{}); r.ctl_exit = r.ctl_entry;
    r.ctl.current_sync = 1 % nsyncs; r.ctl.resync_phase = 0;
    VitArgs a{};
    a.symbols = symbols.data(); a.nchunks = nchunks;
    a.bits_in = tr.bits_in; a.bits_out = tr.bits_out; a.bps = bps; a.nshifts = nsh; a.nsyncs = nsyncs; a.ncs = ncs; a.nsymbols = cst.nsymbols;
    a.path_nbits = tr.path_nbits; a.path_depth = tr.path_depth; a.path32 = tr.path32 ? 1 : 0; a.resync_period = P;
    a.trellis_pred = tr.pred.data(); a.trellis_us = tr.us.data(); a.maps = maps.data(); a.shifts = shifts.data();
    a.state = r.state.data(); a.ctl = &r.ctl; a.out = r.out.data();
    VitSegArgs sg{};
    sg.seg_start = seg_start.data(); sg.nseg = nseg; sg.list = nullptr; sg.nlist = 0; sg.warm_chunks = P > 1 ? 2 : 5; sg.warm_others = P > 1 ? 2 : 0;
    sg.phase0 = 0; sg.nb = nb; sg.entry = r.entry.data(); sg.exit = r.exit.data(); sg.ctl_entry = r.ctl_entry.data(); sg.ctl_exit = r.ctl_exit.data();
    const size_t ws_smem = (((size_t)128 * ncs + (size_t)128 * nb + 15) & ~(size_t)15) + (size_t)4 * ((size_t)nsyncs * 1536 + 128 * 4 + 64 + 128);
    std::vector<unsigned char> dyn(std::max(smem, ws_smem) + 64);
    emu::g_dyn_smem = dyn.data();
    auto launch = [&](unsigned nblk, const VitSegArgs &s2) {
      if (version == 1) {
        if (mode == "r12" || mode == "ws") emu::launch(nblk, 32 * nsyncs, [&] { v1::k_viterbi<true>(a, s2); });
        else emu::launch(nblk, 32 * nsyncs, [&] { v1::k_viterbi<false>(a, s2); });
      } else {
        if (mode == "ws") emu::launch((nblk + v2::kVitWsWarps - 1) / v2::kVitWsWarps, 32 * v2::kVitWsWarps, [&] { v2::k_viterbi_ws(a, s2, nblk); });
        else if (mode == "r12") emu::launch(nblk, 32 * nsyncs, [&] { v2::k_viterbi<v2::kVitR12>(a, s2); });
        else if (mode == "full") emu::launch(nblk, 32 * nsyncs, [&] { v2::k_viterbi<v2::kVitFull>(a, s2); });
        else emu::launch(nblk, 32 * nsyncs, [&] { v2::k_viterbi<v2::kVitGeneric>(a, s2); });
      }
    };
    launch(nseg, sg);
    // the repair path: segments 1 and 2 again, exactly from their predecessors' exit states
    const std::vector<uint32_t> list = {1, 2};
    VitSegArgs rp = sg; rp.list = list.data(); rp.nlist = 2;
    launch(1, rp);
    rp.list = list.data() + 1;
    launch(1, rp);
  };
  if (vs_oracle) {
    // the current text, one serial segment from the constructor state (metrics and paths 0, hypothesis 0, phase 0)
    Run r;
    r.out.assign(nchunks * 16 * tr.bits_in, 0xee);
    r.entry.assign(nsyncs, VitDecState{}); r.exit = r.entry; r.state.assign(nsyncs, VitDecState{});
    r.ctl_entry.assign(1, VitCtl{}); r.ctl_exit = r.ctl_entry; r.ctl = VitCtl{};
    VitArgs a{};
    a.symbols = symbols.data(); a.nchunks = nchunks;
    a.bits_in = tr.bits_in; a.bits_out = tr.bits_out; a.bps = bps; a.nshifts = nsh; a.nsyncs = nsyncs; a.ncs = ncs; a.nsymbols = cst.nsymbols;
    a.path_nbits = tr.path_nbits; a.path_depth = tr.path_depth; a.path32 = tr.path32 ? 1 : 0; a.resync_period = P;
    a.trellis_pred = tr.pred.data(); a.trellis_us = tr.us.data(); a.maps = maps.data(); a.shifts = shifts.data();
    a.state = r.state.data(); a.ctl = &r.ctl; a.out = r.out.data();
    const std::vector<uint64_t> one = {0, nchunks};
    VitSegArgs sg{};
    sg.seg_start = one.data(); sg.nseg = 1; sg.warm_chunks = 2; sg.warm_others = 2; sg.phase0 = 0; sg.nb = nb; sg.full = full ? 1 : 0;
    sg.entry = r.entry.data(); sg.exit = r.exit.data(); sg.ctl_entry = r.ctl_entry.data(); sg.ctl_exit = r.ctl_exit.data();
    const size_t ws_smem = (((size_t)128 * ncs + (size_t)128 * nb + 15) & ~(size_t)15) + (size_t)4 * ((size_t)nsyncs * 1536 + 128 * 4 + 64 + 128);
    std::vector<unsigned char> dyn(std::max(smem, ws_smem) + 64);
    emu::g_dyn_smem = dyn.data();
    if (mode == "ws") emu::launch(1, 32 * v2::kVitWsWarps, [&] { v2::k_viterbi_ws(a, sg, 1); });
    else if (mode == "r12") emu::launch(1, 32 * nsyncs, [&] { v2::k_viterbi<v2::kVitR12>(a, sg); });
    else if (mode == "full") emu::launch(1, 32 * nsyncs, [&] { v2::k_viterbi<v2::kVitFull>(a, sg); });
    else emu::launch(1, 32 * nsyncs, [&] { v2::k_viterbi<v2::kVitGeneric>(a, sg); });
    // the oracle on the same softsymbols (4 bytes each: cost in the low half, symbol above it: sdr.h:455-458)
    static orc_cstln oc;
    if (orc_cstln_build2(&oc, ORC_QPSK, fec, 0)) { fprintf(stderr, "oracle: constellation\n"); return 2; }
    orc_viterbi *ov = orc_viterbi_new(&oc, fec);
    orc_viterbi_set_resync_period(ov, P);
    CHECK(orc_viterbi_nsyncs(ov) == nsyncs && orc_viterbi_nshifts(ov) == nsh && orc_viterbi_bits_in(ov) == tr.bits_in, "oracle configuration");
    std::vector<uint8_t> want(r.out.size(), 0xee);
    size_t consumed = 0;
    const size_t nw = orc_viterbi_run(ov, reinterpret_cast<const uint8_t *>(symbols.data()), nblocks * nsh + (nsh - 1), want.data(), want.size(), &consumed);
    CHECK(nw == want.size() && consumed == nblocks * nsh, "oracle decoded %zu bytes from %zu symbols", nw, consumed);
    CHECK(r.out == want, "output bytes differ from the oracle");
    for (size_t i = 0; i < want.size() && g_fail < 5; ++i) CHECK(r.out[i] == want[i], "out[%zu] (chunk %zu): %02x vs oracle %02x", i, i / (16 * tr.bits_in), r.out[i], want[i]);
    CHECK(r.ctl_exit[0].current_sync == orc_viterbi_current_sync(ov), "elected hypothesis %d vs oracle %d", r.ctl_exit[0].current_sync, orc_viterbi_current_sync(ov));
    CHECK(r.ctl_exit[0].resync_phase == orc_viterbi_resync_phase(ov), "re-sync phase");
    for (int d = 0; d < nsyncs; ++d) {
      int32_t cost[64]; uint64_t path[64];
      orc_viterbi_get_dec(ov, d, cost, path);
      for (int st = 0; st < 64; ++st) CHECK(r.exit[d].cost[st] == cost[st] && r.exit[d].path[st] == path[st], "decoder %d state %d at the end", d, st);
    }
    if (!noise_only) CHECK(r.ctl_exit[0].current_sync == 0, "hypothesis 0 was not elected");
    orc_viterbi_free(ov);
    if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
    printf("identical to the oracle (fec %d, %s, %d decoders, %llu chunks, resync_period %d, current %d)\n", fec, mode.c_str(), nsyncs,
           (unsigned long long)nchunks, P, r.ctl_exit[0].current_sync);
    return 0;
  }
  Run r1, r2;
  run(1, r1);
  run(2, r2);
  CHECK(r1.out == r2.out, "output bytes");
  for (size_t i = 0; i < r1.out.size() && g_fail < 5; ++i) CHECK(r1.out[i] == r2.out[i], "out[%zu] (chunk %zu): %02x vs %02x", i, i / (16 * tr.bits_in), r1.out[i], r2.out[i]);
  auto same_states = [&](const std::vector<VitDecState> &x, const std::vector<VitDecState> &y, const char *what) {
    for (size_t i = 0; i < x.size(); ++i)
      for (int s = 0; s < 64; ++s) CHECK(x[i].cost[s] == y[i].cost[s] && x[i].path[s] == y[i].path[s], "%s[%zu] state %d", what, i, s);
  };
  same_states(r1.entry, r2.entry, "entry");
  same_states(r1.exit, r2.exit, "exit");
  for (uint32_t g = 0; g < nseg; ++g) {
    CHECK(r1.ctl_exit[g].current_sync == r2.ctl_exit[g].current_sync && r1.ctl_exit[g].resync_phase == r2.ctl_exit[g].resync_phase, "ctl_exit[%u]", g);
    if (g) CHECK(r1.ctl_entry[g].current_sync == r2.ctl_entry[g].current_sync, "ctl_entry[%u]", g);
  }
  // the stream must have exercised what it is meant to: some output, and (signal case) hypothesis 0 elected
  size_t written = 0; for (auto b : r2.out) written += b != 0xee;
  CHECK(written > r2.out.size() / 2, "hardly any output written");
  if (!noise_only) CHECK(r2.ctl_exit[nseg - 1].current_sync == 0, "hypothesis 0 was not elected (current %d)", r2.ctl_exit[nseg - 1].current_sync);
  if (g_fail) { fprintf(stderr, "%d mismatches\n", g_fail); return 1; }
  printf("identical (fec %d, %s, %d decoders, %llu chunks, layout %d, current %d)\n", fec, mode.c_str(), nsyncs, (unsigned long long)nchunks, layout, r2.ctl_exit[nseg - 1].current_sync);
  return 0;
}
