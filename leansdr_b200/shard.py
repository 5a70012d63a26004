"""Time-sharded operation: ONE IQ stream over several receivers.

The stream is cut into consecutive chunks; chunk i is handled by rank i % world.  Per chunk:

  1. halo     the last `halo` samples of chunk i-1 ("chunk-edge samples") arrive from the
              neighbour and sit in front of the chunk in HBM;
  2. detect   auto_notch detection on the chunk; the notch bins in force at the chunk end go
              to the next rank (16 bytes);
  3. front    notch + front end + receiver spans, speculative, all ranks concurrently;
  4. back     wait for the previous chunk's EDGE (loop state, seam log, deconvolver /
              sync / de-interleaver / PRBS carry: ldvb_edge_size() bytes), verify and stitch
              the seam, run the exact FEC back end, send this chunk's EDGE on.

Only (4) is serial across ranks.  The transport is torch.distributed point-to-point
(NCCL over NVLink on GPUs; gloo in the CPU tests, which drive this module with a stand-in
engine).  Nothing here touches sample data on the host.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Chunk:
    index: int
    abs_raw0: int      # absolute index of the first sample held (start of the halo)
    n_halo: int
    n_chunk: int
    n_halo_next: int
    last: bool

    @property
    def start(self) -> int:          # absolute index of the first owned sample
        return self.abs_raw0 + self.n_halo


def plan_stream(total_samples: int, n_chunks: int, unit: int, halo: int) -> list[Chunk]:
    """Equal chunks of a multiple of `unit` samples; every chunk but the first gets `halo`
    samples of its predecessor in front.  Samples beyond n_chunks * chunk are not used."""
    if halo % unit:
        raise ValueError("halo must be a multiple of the alignment unit")
    c = total_samples // n_chunks // unit * unit
    if c <= 0:
        raise ValueError("stream too short for this many chunks")
    out = []
    for k in range(n_chunks):
        h = halo if k else 0
        out.append(Chunk(k, k * c - h, h, c, halo if k + 1 < n_chunks else 0, k + 1 == n_chunks))
    return out


class GpuEngine:
    """One receiver handle working on chunks that live in a device buffer."""

    def __init__(self, rx, ts_ptr: int, ts_cap: int):
        self.rx = rx
        self.ts_ptr, self.ts_cap = ts_ptr, ts_cap
        self.edge_size = rx.edge_size()
        self._shard = None

    def detect(self, chunk: Chunk, iq_ptr: int, bins_before):
        self._shard = self.rx.shard(iq_ptr, chunk.abs_raw0, chunk.n_halo, chunk.n_chunk, chunk.n_halo_next, chunk.last,
                                    bins_before)
        return self.rx.shard_detect(self._shard)

    def front(self):
        self.rx.shard_front(self._shard)

    def back(self, edge_in, want_edge: bool):
        edge_out = np.zeros(self.edge_size, np.uint8) if want_edge else None
        npk = self.rx.shard_back(edge_in, self.ts_ptr, self.ts_cap, edge_out)
        return npk, edge_out


def run_local(engines, chunks, iq_ptr_of):
    """Single process: chunk i on engines[i % len(engines)], in stream order.  Returns the
    packets produced per chunk.  iq_ptr_of(chunk) -> device pointer of the chunk's halo."""
    bins = (-1, -1, -1, -1)
    edge = None
    out = []
    for ch in chunks:
        e = engines[ch.index % len(engines)]
        bins_next = e.detect(ch, iq_ptr_of(ch), bins)
        e.front()
        npk, edge = e.back(edge, not ch.last)
        out.append(npk)
        bins = bins_next
    return out


class Ring:
    """Point-to-point transport of the carry messages between neighbouring ranks.

    Messages between a pair of ranks are matched in issue order, so the protocol fixes it: per
    chunk the sender issues halo, bins ("early" group) and EDGE ("edge" group); the receiver
    posts its receives in the same order.  A rank receives its early messages BEFORE it sends
    its own, and early sends are not waited for: when the stream is longer than one round, the
    last rank's messages to rank 0 stay pending until rank 0 gets there, without stalling
    anything else (the EDGE traffic has its own group, hence its own NCCL stream)."""

    def __init__(self, dist, device, separate_groups: bool = True):
        import torch
        self.dist, self.torch, self.device = dist, torch, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.early = dist.new_group() if separate_groups else None
        self.edge = dist.new_group() if separate_groups else None
        self._pending = []

    def flush(self):
        for w, _keep in self._pending:
            w.wait()
        self._pending = []

    def isend_early(self, tensor, dst: int):
        self._pending.append((self.dist.isend(tensor, dst, group=self.early), tensor))

    def isend_early_bytes(self, arr: np.ndarray, dst: int):
        self.isend_early(self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).to(self.device), dst)

    def recv_early(self, tensor, src: int):
        self.dist.recv(tensor, src, group=self.early)

    def recv_early_bytes(self, nbytes: int, src: int) -> np.ndarray:
        t = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)
        self.dist.recv(t, src, group=self.early)
        return t.cpu().numpy()

    def send_edge(self, arr: np.ndarray, dst: int):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8)).to(self.device)
        self.dist.send(t, dst, group=self.edge)

    def recv_edge(self, nbytes: int, src: int) -> np.ndarray:
        t = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)
        self.dist.recv(t, src, group=self.edge)
        return t.cpu().numpy()


def run_round(engine, ring: Ring, chunk: Chunk, iq_ptr: int, halo_send=None, halo_recv=None, timeline=None):
    """This rank's turn in the ring: `chunk` is chunk number chunk.index of the stream and
    chunk.index % ring.world == ring.rank.  halo_send: device view of this chunk's last
    n_halo_next samples; halo_recv: device view of the halo region in front of the chunk.
    Returns the number of TS packets produced."""
    import time
    r, n = ring.rank, ring.world
    first, last = chunk.index == 0, chunk.last
    prv, nxt = (r - 1) % n, (r + 1) % n
    t0 = time.perf_counter()
    ring.flush()
    bins = (-1, -1, -1, -1)
    if not first:
        ring.recv_early(halo_recv, prv)
        bins = tuple(int(v) for v in ring.recv_early_bytes(16, prv).view(np.int32))
    after = engine.detect(chunk, iq_ptr, bins)
    if not last:
        ring.isend_early(halo_send, nxt)
        ring.isend_early_bytes(np.asarray(after, np.int32), nxt)
    t1 = time.perf_counter()
    engine.front()
    t2 = time.perf_counter()
    edge_in = None if first else ring.recv_edge(engine.edge_size, prv)
    t3 = time.perf_counter()
    npk, edge_out = engine.back(edge_in, not last)
    if not last:
        ring.send_edge(edge_out, nxt)
    t4 = time.perf_counter()
    if timeline is not None:   # host wall clock per phase, ms (accumulated)
        for k, v in (("early", t1 - t0), ("front", t2 - t1), ("wait_edge", t3 - t2), ("back", t4 - t3)):
            timeline[k] = timeline.get(k, 0.0) + v * 1e3
    return npk
