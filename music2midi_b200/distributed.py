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


def gather_tokens(local_tokens: torch.Tensor, n_total: int, group=None, counts: Optional[List[int]] = None,
                  out_dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """All-gather of per-rank token blocks [n_local, L] (any integer dtype) into [n_total, L] in global segment
    order.  ONE collective when `counts` (rows held by every rank) is known or every rank holds n_total / world
    rows; otherwise a second, tiny one exchanges the counts first.  Token ids travel as int16 (vocabulary 400) and
    are only widened if `out_dtype` asks for it."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_tokens.to(out_dtype)
    world = dist.get_world_size(group)
    L = local_tokens.shape[1]
    if counts is None and n_total == world * local_tokens.shape[0]:
        # contiguous block sharding differs by at most one clip between ranks: if this rank holds exactly
        # n_total / world rows, every rank does
        counts = [local_tokens.shape[0]] * world
    if counts is None:
        n_local = torch.tensor([local_tokens.shape[0]], dtype=torch.int64, device=local_tokens.device)
        got = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(got, n_local, group=group)
        counts = [int(c) for c in got]
    counts = [int(c) for c in counts]
    if len(counts) != world or sum(counts) != n_total:
        raise RuntimeError(f"token gather: ranks hold {sum(counts)} rows, expected {n_total}")
    if counts[dist.get_rank(group)] != local_tokens.shape[0]:
        raise RuntimeError("token gather: counts do not match this rank's block")
    width = max(counts)
    if local_tokens.dtype == torch.int16 and local_tokens.shape[0] == width and local_tokens.is_contiguous():
        send = local_tokens
    else:
        send = torch.zeros(width, L, dtype=torch.int16, device=local_tokens.device)
        send[: local_tokens.shape[0]] = local_tokens.to(torch.int16)
    recv = torch.empty(world * width, L, dtype=torch.int16, device=local_tokens.device)
    # transported as raw bytes: gloo (CPU tests) has no int16 collectives, NCCL does not care
    dist.all_gather_into_tensor(recv.view(torch.uint8), send.view(torch.uint8), group=group)
    if all(c == width for c in counts):
        out = recv
    else:
        out = torch.cat([recv[r * width: r * width + counts[r]] for r in range(world)], dim=0)
    return out if out_dtype == torch.int16 else out.to(out_dtype)


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
