"""Host side of the time-sharded mode (leansdr_b200/shard.py) on CPU: chunk planning and the
ring protocol (halo exchange, notch-bin chain, EDGE chain) with world_size 2 over gloo.  A
stand-in engine replaces the CUDA handle: its EDGE is a running SHA-256 over everything the
chunk's owner must have seen (halo received from the neighbour, bins, previous EDGE), so the
final digest equals a serial computation only if every message went to the right place in the
right order."""
import hashlib
import os
import socket

import numpy as np
import pytest

from leansdr_b200 import shard as S


def test_plan_stream_geometry():
    ch = S.plan_stream(10_000_000, 4, 4096, 8192)
    assert [c.index for c in ch] == [0, 1, 2, 3]
    c = ch[0].n_chunk
    assert c % 4096 == 0 and 4 * c <= 10_000_000 < 4 * (c + 4096)
    assert ch[0].abs_raw0 == 0 and ch[0].n_halo == 0 and not ch[0].last
    for k in (1, 2, 3):
        assert ch[k].n_halo == 8192 and ch[k].abs_raw0 == k * c - 8192 and ch[k].start == k * c
        assert ch[k - 1].n_halo_next == ch[k].n_halo
    assert ch[3].last and ch[3].n_halo_next == 0
    with pytest.raises(ValueError):
        S.plan_stream(1000, 4, 4096, 4096)
    with pytest.raises(ValueError):
        S.plan_stream(1 << 20, 2, 4096, 100)


class HashEngine:
    edge_size = 32

    def __init__(self, buf):
        self.buf = buf            # torch uint8 [halo | chunk]

    def detect(self, chunk, iq_ptr, bins):
        self.chunk, self.bins = chunk, bins
        return tuple(b + chunk.index + 1 for b in bins)

    def front(self):
        pass

    def back(self, edge_in, want_edge):
        h = hashlib.sha256()
        h.update(b"" if edge_in is None else edge_in.tobytes())
        h.update(np.asarray(self.bins, np.int32).tobytes())
        n = self.chunk.n_halo + self.chunk.n_chunk
        off = self.buf.numel() - self.chunk.n_chunk - self.chunk.n_halo
        h.update(self.buf.numpy()[off: off + n].tobytes())
        self.digest = h.digest()
        return self.chunk.index, (np.frombuffer(self.digest, np.uint8).copy() if want_edge else None)


def serial_digest(stream, chunks):
    edge, bins = b"", (-1, -1, -1, -1)
    for c in chunks:
        h = hashlib.sha256()
        h.update(edge)
        h.update(np.asarray(bins, np.int32).tobytes())
        h.update(stream[c.abs_raw0: c.abs_raw0 + c.n_halo + c.n_chunk].tobytes())
        edge = h.digest()
        bins = tuple(b + c.index + 1 for b in bins)
    return edge


def _worker(rank, world, port, rounds, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        halo, total = 64, 4096
        stream = np.random.default_rng(5).integers(0, 256, total, dtype=np.uint8)
        chunks = S.plan_stream(total, world * rounds, 16, halo)
        ring = S.Ring(dist, torch.device("cpu"))
        got = []
        for j in range(rounds):
            ch = chunks[j * world + rank]
            buf = torch.zeros(halo + ch.n_chunk, dtype=torch.uint8)
            buf[halo:] = torch.from_numpy(stream[ch.start: ch.start + ch.n_chunk])   # the chunk is resident,
            eng = HashEngine(buf)                                                    # the halo is not
            npk = S.run_round(eng, ring, ch, 0, halo_send=buf[buf.numel() - halo:],
                              halo_recv=buf[:halo])
            assert npk == ch.index
            got.append(eng.digest)
        if chunks[-1].index % world == rank:
            q.put(("last", got[-1], serial_digest(stream, chunks)))
        ring.flush()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rounds", [1, 3])
def test_ring_protocol_world2_gloo(rounds):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rounds, q)) for r in range(2)]
    for p in procs:
        p.start()
    tag, got, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert tag == "last" and got == want
