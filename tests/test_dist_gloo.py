"""world_size-2 (and 3) gloo tests of the data-parallel sharding + caption all-gather
(on_device_image_captioning_b200/dist.py).  The per-rank captioner is a deterministic stand-in
so the test needs no GPU; the collective and the index arithmetic are the code under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_caption(x):
    # "caption" = a function of the image content only, so order mistakes are visible
    n = x.shape[0]
    key = x.reshape(n, -1).sum(dim=1).round().to(torch.int32)
    L = 6
    tok = torch.stack([key + i for i in range(L)], dim=1).reshape(n, 1, L).to(torch.int32)
    ln = (key % L + 1).reshape(n, 1).to(torch.int32)
    return tok, ln


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from on_device_image_captioning_b200 import dist as D
    x = torch.arange(n_items, dtype=torch.float32).reshape(n_items, 1, 1, 1) * 10.0
    tok, ln = D.caption_sharded(_fake_caption, x, rank, world)
    ref_tok, ref_ln = _fake_caption(x)
    ok = bool(torch.equal(tok, ref_tok) and torch.equal(ln, ref_ln))
    q.put((rank, ok, tuple(tok.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 8), (2, 7), (3, 5), (2, 1)])
def test_sharded_caption_gather(world, n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape[0] == n_items for _, _, shape in res), res


def test_shard_indices_partition():
    from on_device_image_captioning_b200 import dist as D
    for n in (0, 1, 7, 64):
        for w in (1, 2, 3, 8):
            seen = sorted(i for r in range(w) for i in D.shard_indices(n, r, w))
            assert seen == list(range(n))
