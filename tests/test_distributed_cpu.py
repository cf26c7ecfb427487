"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: clip-wise sharding and the token gather."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from music2midi_b200.distributed import gather_tokens, shard_clips, shard_range, transcribe_sharded


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 2048, 2049):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard_clips(5, 10, 1, 2) == (30, 50)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_generate(wave, cond):
    # deterministic stand-in for the GPU hot path: tokens depend only on the row's own data
    base = (wave.abs().sum(1) * 1000).long() % 397
    L = 6
    t = (base[:, None] + torch.arange(L)[None, :] * (1 + cond[:, :1])) % 400
    t[:, 0] = 1
    return t


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        segs = torch.randn(n_clips * 10, 32, generator=g)
        cond = torch.randint(0, 3, (n_clips * 10, 2), generator=g)
        out = transcribe_sharded(_fake_generate, segs, cond, 10, max_length=8)
        lo, hi = shard_clips(n_clips, 10, rank, world)
        local = torch.zeros(hi - lo, 8, dtype=torch.int16)
        local[:, :6] = _fake_generate(segs[lo:hi], cond[lo:hi]).to(torch.int16)
        out2 = gather_tokens(local, n_clips * 10)
        q.put((rank, out.tolist(), out2.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [4, 5])
def test_two_rank_gloo_matches_single_process(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_clips
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    segs = torch.randn(n_clips * 10, 32, generator=g)
    cond = torch.randint(0, 3, (n_clips * 10, 2), generator=g)
    expect = torch.zeros(n_clips * 10, 8, dtype=torch.int64)
    expect[:, :6] = _fake_generate(segs, cond)
    for rank, out, out2 in res:
        assert out == expect.tolist() and out2 == expect.tolist(), rank


def test_single_process_path_needs_no_process_group():
    t = torch.arange(12, dtype=torch.int16).reshape(3, 4)
    assert torch.equal(gather_tokens(t, 3), t.to(torch.int64))
