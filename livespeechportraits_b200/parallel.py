"""Frame-sharded rendering across GPUs: one process per GPU, weights replicated, frames partitioned.

The reference's only multi-GPU hook is ``nn.DataParallel`` (models/networks.py:392-401), single process.
Frames are independent (eval-mode BatchNorm, no temporal state: demo.py:260-266), so the B200 layout is:
contiguous block partition of the clip's frames over the ranks, no data-path collective while rendering,
and ONE exchange step - an all-gather of the rendered frames - chunked so that the gather of chunk *i* runs
on a side stream over NVLink while chunk *i+1* is being rendered.  The tail kernel writes each rank's frames
straight into its slot of the (in-place) gather buffer, so there is no staging copy before the collective.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def partition(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced block partition: the first ``n_items % world`` ranks get one extra item."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def chunk_schedule(n_local_max: int, chunk: int) -> List[Tuple[int, int]]:
    """[(offset, length)] over the LONGEST shard; shorter shards render fewer frames in the last chunk(s)."""
    if chunk < 1:
        raise ValueError("chunk must be >= 1")
    return [(o, min(chunk, n_local_max - o)) for o in range(0, n_local_max, chunk)]


class ShardedRenderer:
    """Renders this rank's shard chunk by chunk and all-gathers the frames of every rank.

    ``render_fn(feature_maps[n,1,H,W], out[n,3,H,W]) -> None`` renders into ``out`` in place (on GPU this is
    ``Feature2Face_G.render(fm, cand, out=out)``; tests on CPU/gloo pass a torch function).
    """

    def __init__(self, render_fn: Callable[[torch.Tensor, torch.Tensor], None], group=None, chunk: int = 48):
        self.render_fn = render_fn
        self.group = group
        self.chunk = chunk
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def render(self, n_total: int, local_feature_maps: torch.Tensor, gather: bool = True) -> torch.Tensor:
        """``local_feature_maps``: this rank's block ``partition(n_total, world, rank)`` of the clip.

        Returns ``[n_total,3,H,W]`` (every rank holds all frames, frame order = clip order) when ``gather``,
        else this rank's ``[n_local,3,H,W]``.
        """
        start, stop = partition(n_total, self.world, self.rank)
        n_local = stop - start
        if local_feature_maps.shape[0] != n_local:
            raise ValueError(f"rank {self.rank} expected {n_local} frames, got {local_feature_maps.shape[0]}")
        dev = local_feature_maps.device
        h, w = local_feature_maps.shape[-2:]
        if not gather or self.world == 1:
            out = torch.empty((n_local, 3, h, w), dtype=torch.float32, device=dev)
            for off, ln in chunk_schedule(n_local, self.chunk):
                self.render_fn(local_feature_maps[off:off + ln], out[off:off + ln])
            return out
        n_max = partition(n_total, self.world, 0)[1]          # rank 0 holds the longest shard
        final = torch.empty((n_total, 3, h, w), dtype=torch.float32, device=dev)
        sched = chunk_schedule(n_max, self.chunk)
        bufs = [torch.empty((self.world, self.chunk, 3, h, w), dtype=torch.float32, device=dev) for _ in range(2)]
        use_streams = dev.type == "cuda"
        comm = torch.cuda.Stream(device=dev) if use_streams else None
        if use_streams:                      # tensors allocated on the compute stream, also used on the comm stream
            final.record_stream(comm)
            for b_ in bufs:
                b_.record_stream(comm)
        done: List[Optional[torch.cuda.Event]] = [None, None]
        bounds = [partition(n_total, self.world, r) for r in range(self.world)]
        for ci, (off, ln) in enumerate(sched):
            buf = bufs[ci & 1]
            if use_streams and done[ci & 1] is not None:
                torch.cuda.current_stream(dev).wait_event(done[ci & 1])   # buffer free again
            mine = max(0, min(ln, n_local - off))
            if mine > 0:
                self.render_fn(local_feature_maps[off:off + mine], buf[self.rank, :mine])
            if use_streams:
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(dev))
                comm.wait_event(ready)
                ctx = torch.cuda.stream(comm)
            else:
                import contextlib
                ctx = contextlib.nullcontext()
            with ctx:
                flat = buf.view(self.world * self.chunk, 3, h, w)
                dist.all_gather_into_tensor(flat, buf[self.rank], group=self.group)   # in place: slot = rank
                for r, (s, e) in enumerate(bounds):
                    cnt = max(0, min(ln, (e - s) - off))
                    if cnt > 0:
                        final[s + off:s + off + cnt].copy_(buf[r, :cnt], non_blocking=True)
                if use_streams:
                    ev = torch.cuda.Event()
                    ev.record(comm)
                    done[ci & 1] = ev
        if use_streams:
            torch.cuda.current_stream(dev).wait_stream(comm)
        return final
