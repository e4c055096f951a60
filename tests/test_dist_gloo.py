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


class _FakeSwinEngine:
    """Stand-in for Engine on the feature-extraction path (preprocess_rgb8 + forward_swin): features are a function of
    the image content only, so a mix-up of ids, ranks or batches is visible."""

    def preprocess_rgb8(self, images, img_size=None):
        return torch.stack([torch.as_tensor(im, dtype=torch.float32).mean().reshape(1) for im in images])

    def forward_swin(self, x):
        return x.reshape(-1, 1, 1) + torch.arange(6, dtype=torch.float32).reshape(1, 3, 2)


def _feature_worker(rank, world, port, n_items, out_path, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from on_device_image_captioning_b200 import features as F
    images = [np.full((4, 5, 3), i, dtype=np.uint8) for i in range(n_items)]
    ids = [1000 + i for i in range(n_items)]
    path = F.extract_features_sharded(_FakeSwinEngine(), images, ids, out_path, batch_size=2)     # rank / world from the group
    q.put((rank, path))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 7), (3, 4)])
def test_feature_extraction_sharded_by_rank(world, n_items, tmp_path):
    """SURVEY.md 8f N3 across ranks: every rank writes its own container, the union holds every image exactly once under
    the reference's "<img_id>_features" keys, and each entry is the feature of ITS image."""
    import numpy as np
    from on_device_image_captioning_b200 import features as F
    out = str(tmp_path / "precalc_features.hdf5")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_feature_worker, args=(r, world, port, n_items, out, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert sorted(res.values()) == sorted(F.shard_path(out, r, world) for r in range(world))
    shards = F.FeatureShards(out, world)
    assert len(shards) == n_items
    for i in range(n_items):
        assert (1000 + i) in shards
        want = float(i) + np.arange(6, dtype=np.float32).reshape(3, 2)
        np.testing.assert_array_equal(shards.read(1000 + i), want)
    assert 999 not in shards
    assert F.shard_path(out, 0, 1) == out
