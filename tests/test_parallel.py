"""Frame sharding + chunked all-gather, exercised on CPU with the gloo backend (world_size 2 and 3)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from livespeechportraits_b200.parallel import ShardedRenderer, chunk_schedule, partition


def test_partition_is_a_contiguous_cover():
    for n in (0, 1, 7, 8, 11, 10000):
        for world in (1, 2, 3, 8):
            edges = [partition(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert partition(10000, 8, 3) == (3750, 5000)
    with pytest.raises(ValueError):
        partition(4, 2, 2)


def test_chunk_schedule():
    assert chunk_schedule(7, 3) == [(0, 3), (3, 3), (6, 1)]
    assert chunk_schedule(0, 3) == []


def _fake_render(fm, out):
    # deterministic per-frame function standing in for the generator: frames must end up in clip order
    m = fm.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1)
    out.copy_(m * torch.tensor([1.0, 2.0, 3.0]).view(1, 3, 1, 1) + fm)


def _fake_render_u8(fm, out):
    # uint8 HWC stand-in for Feature2Face_G.render_image
    m = (fm.mean(dim=(1, 2, 3)) * 200).to(torch.uint8).view(-1, 1, 1, 1)
    out.copy_(m + (fm[:, 0, :, :, None] * 50).to(torch.uint8) + torch.tensor([0, 1, 2], dtype=torch.uint8).view(1, 1, 1, 3))


def _worker(rank, world, port, n_total, chunk, gather, uint8, to_host, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        clip = torch.rand(n_total, 1, 8, 8, generator=g)
        s, e = partition(n_total, world, rank)
        fn = _fake_render_u8 if uint8 else _fake_render
        r = ShardedRenderer(fn, chunk=chunk, uint8=uint8)
        exp = torch.empty((n_total, 8, 8, 3), dtype=torch.uint8) if uint8 else torch.empty(n_total, 3, 8, 8)
        fn(clip, exp)
        host = torch.zeros_like(exp) if to_host else None
        out = r.render(n_total, clip[s:e].clone(), gather=gather, host_out=host if rank == 0 else None, to_host=to_host)
        if gather:
            ok = torch.equal(out, exp) and r.gather_mode == "nccl"       # CPU tensors: the collective path (gloo here)
            if to_host and rank == 0:
                ok = ok and torch.equal(host, exp)                       # frames delivered to the host in clip order
        else:
            ok = torch.equal(out, exp[s:e])
        again = r.render(n_total, clip[s:e].clone(), gather=gather)      # the renderer is reusable
        ok = ok and torch.equal(again, exp if gather else exp[s:e])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_total,chunk,gather,uint8,to_host", [
    (2, 11, 3, True, False, False), (3, 10, 4, True, False, True), (2, 5, 8, True, True, True), (2, 6, 2, False, False, False),
    (2, 1, 4, True, True, False),       # fewer frames than ranks: rank 1 renders nothing but still takes part in the gather
])
def test_sharded_render_gloo(world, n_total, chunk, gather, uint8, to_host):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, chunk, gather, uint8, to_host, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = sorted(q.get(timeout=5) for _ in range(world))
    assert results == [(r, True) for r in range(world)]


def test_gather_mode_argument_is_validated():
    with pytest.raises(ValueError):
        ShardedRenderer(_fake_render, gather="smoke-signals")
    r = ShardedRenderer(_fake_render, gather="ce")          # single process: world 1, nothing to gather
    out = r.render(3, torch.rand(3, 1, 8, 8))
    assert out.shape == (3, 3, 8, 8) and r.gather_mode is None
