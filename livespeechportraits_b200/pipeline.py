"""Batched render loop: host feature maps in, host frames out, copies overlapped with the generator.

The reference renders one frame per Python iteration with a blocking ``.to(device)`` before and a blocking
``.cpu()`` after every frame (demo.py:260-272, util/util.py:33).  Every feature map of a clip is known before
the loop starts, so here the clip is cut into batches that flow through three CUDA streams
(H2D copy -> generator -> D2H copy) with double-buffered device staging: the PCIe copies of batch *i+1* and
*i-1* overlap the kernels of batch *i*.  Host buffers must be pinned for the copies to be asynchronous.
"""
from __future__ import annotations

from typing import Optional

import torch


class ClipRenderer:
    def __init__(self, net, batch: int = 8, device: Optional[torch.device] = None, precision: Optional[str] = None,
                 uint8: bool = False):
        """``uint8=True``: frames leave the GPU as uint8 HWC images (util.tensor2im fused into the last kernel)."""
        self.net = net
        self.uint8 = uint8
        self.batch = int(batch)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self._h2d = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)
        self._compute = torch.cuda.Stream(self.device)
        self._fm = None
        self._out = None

    def _staging(self, h: int, w: int):
        if self._fm is None or self._fm[0].shape[-2:] != (h, w):
            self._fm = [torch.empty((self.batch, 1, h, w), dtype=torch.float32, device=self.device) for _ in range(2)]
            oshape, odt = ((self.batch, h, w, 3), torch.uint8) if self.uint8 else ((self.batch, 3, h, w), torch.float32)
            self._out = [torch.empty(oshape, dtype=odt, device=self.device) for _ in range(2)]
        return self._fm, self._out

    def render_clip_from_landmarks(self, landmarks_host: torch.Tensor, shoulders_host: Optional[torch.Tensor],
                                   cand_device: torch.Tensor, out_host: torch.Tensor, size: tuple = (512, 512),
                                   on_batch=None) -> torch.Tensor:
        """Same loop with the feature maps drawn on the GPU (``Feature2Face_G.draw_feature_maps``, the batched replacement of
        datasets/face_dataset.py:276-323): per frame 728 bytes of landmark / shoulder tracks cross PCIe instead of a 1 MB map.
        ``landmarks_host`` [N,73,2] and ``shoulders_host`` [N,2k,2] (or None) are fp32; the whole clip's tracks (a few hundred
        KB) are uploaded once, batches are rasterised on the compute stream right before their generator pass.
        ``on_batch(offset, length)`` (optional) is called on the host as soon as the frames ``out_host[offset:offset+length]``
        have landed, in clip order, while later batches are still rendering (the video writer of video.py hangs here)."""
        n = landmarks_host.shape[0]
        w, h = int(size[0]), int(size[1])
        if tuple(out_host.shape) != ((n, h, w, 3) if self.uint8 else (n, 3, h, w)) or not out_host.is_pinned():
            raise ValueError("out_host must be pinned [N,3,H,W] fp32 (or [N,H,W,3] uint8 with uint8=True)")
        fm_dev, out_dev = self._staging(h, w)
        caller = torch.cuda.current_stream(self.device)
        for s in (self._compute, self._d2h):
            s.wait_stream(caller)
        per_frame_cand = cand_device.shape[0] == n and n > 1
        d2h_done = [None, None]
        landed = []
        with torch.cuda.stream(self._compute):
            lm_dev = landmarks_host.to(self.device, non_blocking=True)
            sh_dev = shoulders_host.to(self.device, non_blocking=True) if shoulders_host is not None else None
        for bi, off in enumerate(range(0, n, self.batch)):
            ln = min(self.batch, n - off)
            j = bi & 1
            with torch.cuda.stream(self._compute):
                if d2h_done[j] is not None:
                    self._compute.wait_event(d2h_done[j])        # previous frames of this buffer are on the host
                self.net.draw_feature_maps(lm_dev[off:off + ln], None if sh_dev is None else sh_dev[off:off + ln], (w, h),
                                           out=fm_dev[j][:ln])
                cd = cand_device[off:off + ln] if per_frame_cand else cand_device[:1]
                self.net.render(fm_dev[j][:ln], cd, out=out_dev[j][:ln], precision=self.precision, _uint8=self.uint8)
                ev = torch.cuda.Event()
                ev.record(self._compute)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(ev)
                out_host[off:off + ln].copy_(out_dev[j][:ln], non_blocking=True)
                dn = torch.cuda.Event()
                dn.record(self._d2h)
                d2h_done[j] = dn
                landed.append((off, ln, dn))
        if on_batch is not None:
            for off, ln, dn in landed:
                dn.synchronize()
                on_batch(off, ln)
        self._d2h.synchronize()
        caller.wait_stream(self._compute)
        return out_host

    def render_clip(self, feature_maps_host: torch.Tensor, cand_device: torch.Tensor, out_host: torch.Tensor) -> torch.Tensor:
        """``feature_maps_host`` [N,1,H,W] fp32 (pinned), ``cand_device`` [1|N,12,H,W] on the GPU (demo.py:95 moves the
        candidates once per clip), ``out_host`` [N,3,H,W] fp32 (pinned).  Returns ``out_host`` after the last copy
        has landed (the only host synchronisation of the call)."""
        n, _, h, w = feature_maps_host.shape
        if tuple(out_host.shape) != ((n, h, w, 3) if self.uint8 else (n, 3, h, w)):
            raise ValueError("out_host must be [N,3,H,W] fp32 (or [N,H,W,3] uint8 with uint8=True)")
        if not (feature_maps_host.is_pinned() and out_host.is_pinned()):
            raise ValueError("host buffers must be pinned (torch.empty(..., pin_memory=True)) for asynchronous copies")
        fm_dev, out_dev = self._staging(h, w)
        caller = torch.cuda.current_stream(self.device)
        for s in (self._h2d, self._compute, self._d2h):
            s.wait_stream(caller)
        comp_done = [None, None]
        d2h_done = [None, None]
        per_frame_cand = cand_device.shape[0] == n and n > 1
        for bi, off in enumerate(range(0, n, self.batch)):
            ln = min(self.batch, n - off)
            j = bi & 1
            with torch.cuda.stream(self._h2d):
                if comp_done[j] is not None:
                    self._h2d.wait_event(comp_done[j])          # generator finished reading this staging buffer
                fm_dev[j][:ln].copy_(feature_maps_host[off:off + ln], non_blocking=True)
                up = torch.cuda.Event()
                up.record(self._h2d)
            with torch.cuda.stream(self._compute):
                self._compute.wait_event(up)
                if d2h_done[j] is not None:
                    self._compute.wait_event(d2h_done[j])        # previous frames of this buffer are on the host
                cd = cand_device[off:off + ln] if per_frame_cand else cand_device[:1]
                self.net.render(fm_dev[j][:ln], cd, out=out_dev[j][:ln], precision=self.precision, _uint8=self.uint8)
                ev = torch.cuda.Event()
                ev.record(self._compute)
                comp_done[j] = ev
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(ev)
                out_host[off:off + ln].copy_(out_dev[j][:ln], non_blocking=True)
                dn = torch.cuda.Event()
                dn.record(self._d2h)
                d2h_done[j] = dn
        self._d2h.synchronize()
        caller.wait_stream(self._compute)
        return out_host
