"""CPU suite: the N>1 host path (contiguous image shards + all-gather of captions) on world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from capdec_b200.sharding import gather_captions, shard_range


def test_shard_range_covers_batch():
    for n in (0, 1, 7, 64, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def _fake_decode(start, end, T=6):
    """stand-in for a rank's decode: captions are a deterministic function of the global image index"""
    idx = torch.arange(start, end)
    tokens = (idx[:, None] * 7 + torch.arange(T)[None, :]).int()
    return {"tokens": tokens, "lengths": (idx % T + 1).int(), "scores": -idx.float() / 3}


def _worker(rank, world, port, n_images, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = shard_range(n_images, rank, world)
        full = gather_captions(_fake_decode(s, e), n_images)
        want = _fake_decode(0, n_images)
        good = all(torch.equal(full[k], want[k]) for k in want)
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_gather_captions_world2_gloo_even_and_ragged():
    for n_images in (8, 9):
        ok = mp.get_context("spawn").Array("i", [0, 0])
        mp.spawn(_worker, args=(2, _free_port(), n_images, ok), nprocs=2, join=True)
        assert list(ok) == [1, 1]
