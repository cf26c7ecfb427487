"""Clip-wise sharding over the GPUs of one box (one process per GPU, torch.distributed).

Segments are independent (no model state crosses segments; the tokenizer's sequential mode only
adds 60*segment_index to time indices — reference music2midi/tokenizer.py:75-83, model.py:113-139),
so the data path has NO collective: each rank transcribes a contiguous block of clips with replicated
weights.  The only exchange is one all-gather of the decoded token streams at the end (int16: the
vocabulary has 400 ids), over NCCL/NVLink on GPUs or gloo on CPU (tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; block sizes differ by at most one item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_clips(n_clips: int, segments_per_clip: int, rank: int, world: int) -> Tuple[int, int]:
    """Segment range of this rank when whole clips are kept together (clip-wise partition)."""
    lo, hi = shard_range(n_clips, rank, world)
    return lo * segments_per_clip, hi * segments_per_clip


def gather_tokens(local_tokens: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of per-rank token blocks [n_local, L] (any integer dtype) into [n_total, L] int64 in
    global segment order.  Blocks may differ in size by one clip: ranks pad to the largest block."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_tokens.to(torch.int64)
    world = dist.get_world_size(group)
    L = local_tokens.shape[1]
    n_local = torch.tensor([local_tokens.shape[0]], dtype=torch.int64, device=local_tokens.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c) for c in counts]
    if sum(counts) != n_total:
        raise RuntimeError(f"token gather: ranks hold {sum(counts)} rows, expected {n_total}")
    width = max(counts)
    send = torch.zeros(width, L, dtype=torch.int16, device=local_tokens.device)
    send[: local_tokens.shape[0]] = local_tokens.to(torch.int16)
    recv = torch.empty(world * width, L, dtype=torch.int16, device=local_tokens.device)
    # transported as raw bytes: gloo (CPU tests) has no int16 collectives, NCCL does not care
    dist.all_gather_into_tensor(recv.view(torch.uint8), send.view(torch.uint8), group=group)
    parts = [recv[r * width: r * width + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0).to(torch.int64)


def transcribe_sharded(generate_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], segments: torch.Tensor,
                       cond: torch.Tensor, segments_per_clip: int, max_length: int = 1024,
                       group=None) -> torch.Tensor:
    """Every rank passes the same [n_seg, S] batch description (or just its own rows, see below) and
    receives all token rows [n_seg, max_length] (zero padded).

    `generate_fn(wave[n,S], cond[n,2]) -> tokens[n, <=max_length]` is the single-GPU hot path
    (T5Transformer.generate / Engine.generate)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_seg = segments.shape[0]
    if n_seg % segments_per_clip:
        raise ValueError("segments must be whole clips")
    lo, hi = shard_clips(n_seg // segments_per_clip, segments_per_clip, rank, world)
    tok = generate_fn(segments[lo:hi], cond[lo:hi])
    full = torch.zeros(hi - lo, max_length, dtype=torch.int16, device=tok.device)
    full[:, : tok.shape[1]] = tok.to(torch.int16)
    return gather_tokens(full, n_seg, group)
