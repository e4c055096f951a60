"""Data-parallel captioning over the GPUs of one box.

Captions are independent per image, so the path shards with no data-path collective: rank r
owns images r, r+W, r+2W, ... (weights replicated), and the only exchange is one all-gather of
the int32 caption token ids (+ lengths) at the end of a batch -- B_local x max_len x 4 bytes
per rank.  The reference has no counterpart: its test.py replicates inference on every rank
(test.py:310) and only DDP's gradient all-reduce ever communicates.

`caption_fn` abstracts the per-rank captioner so the sharding/gather logic is testable on CPU
with the gloo backend (tests/test_dist_gloo.py); on GPUs it is Engine.beam_search over NCCL.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


def padded_local_count(n_items: int, world: int) -> int:
    return (n_items + world - 1) // world


def gather_captions(tokens: torch.Tensor, lengths: torch.Tensor, n_items: int, rank: int, world: int,
                    group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """tokens (n_local, how_many, L) int32, lengths (n_local, how_many) int32 for this rank's shard ->
    (n_items, how_many, L), (n_items, how_many) in the original image order, on every rank."""
    if world == 1:
        return tokens, lengths
    n_pad = padded_local_count(n_items, world)
    how_many, L = tokens.shape[1], tokens.shape[2]
    tk = torch.full((n_pad, how_many, L), -1, dtype=torch.int32, device=tokens.device)
    ln = torch.zeros((n_pad, how_many), dtype=torch.int32, device=tokens.device)
    tk[:tokens.shape[0]] = tokens
    ln[:lengths.shape[0]] = lengths
    all_tk = [torch.empty_like(tk) for _ in range(world)]
    all_ln = [torch.empty_like(ln) for _ in range(world)]
    dist.all_gather(all_tk, tk, group=group)
    dist.all_gather(all_ln, ln, group=group)
    out_tk = torch.stack(all_tk, dim=1).reshape(n_pad * world, how_many, L)[:n_items]   # index i*W + r == image id
    out_ln = torch.stack(all_ln, dim=1).reshape(n_pad * world, how_many)[:n_items]
    return out_tk, out_ln


def caption_sharded(caption_fn: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], inputs: torch.Tensor,
                    rank: int, world: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Caption `inputs` (same tensor on every rank) data-parallel: each rank runs `caption_fn` on its
    shard and the token ids are all-gathered."""
    idx = shard_indices(inputs.shape[0], rank, world)
    local = inputs[idx] if idx else inputs[:0]
    if len(idx):
        tok, ln = caption_fn(local)
    else:
        tok = torch.empty(0, 1, 1, dtype=torch.int32, device=inputs.device)
        ln = torch.empty(0, 1, dtype=torch.int32, device=inputs.device)
    if world > 1 and not len(idx):
        # shape agreement for the padded gather: learn (how_many, L) from rank 0
        shp = torch.zeros(2, dtype=torch.int64, device=inputs.device)
        dist.broadcast(shp, src=0, group=group)
        tok = torch.empty(0, int(shp[0]), int(shp[1]), dtype=torch.int32, device=inputs.device)
        ln = torch.empty(0, int(shp[0]), dtype=torch.int32, device=inputs.device)
    elif world > 1 and inputs.shape[0] < world:
        shp = torch.tensor([tok.shape[1], tok.shape[2]], dtype=torch.int64, device=inputs.device)
        dist.broadcast(shp, src=0, group=group)
    return gather_captions(tok, ln, inputs.shape[0], rank, world, group)
