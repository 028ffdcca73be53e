"""Multi-GPU tile sharding: scene replicated per GPU, tile i -> rank i mod N, one
all-gather of equal-sized slot buffers per frame.  The reference has no distributed layer; its only
parallelism is `parMap` over the same 65x65 tiles (Glome.hs:379-386), and tiles never read across
their edges, so the N-GPU frame is bit-identical to the 1-GPU frame.

torch is plumbing here: device memory, streams, and torch.distributed (NCCL over NVLink).
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .scene import render_opts, tile_rects


def tile_slots(width, height, blocksize, world):
    return L.load().glome_tile_slots(width, height, blocksize, world)


def pack_host(frame, blocksize, first, stride):
    """numpy statement of k_tiles_copy<PACK>: frame[h,w,...] -> [slots, bs*bs, ...] (zero padded)."""
    h, w = frame.shape[:2]
    rects = tile_rects(w, h, blocksize)
    slots = (len(rects) + stride - 1) // stride
    out = np.zeros((slots, blocksize * blocksize) + frame.shape[2:], dtype=frame.dtype)
    for k, ti in enumerate(range(first, len(rects), stride)):
        x, y, tw, th = rects[ti]
        out[k, :tw * th] = frame[y:y + th, x:x + tw].reshape((tw * th,) + frame.shape[2:])
    return out


def unpack_host(packed, frame, blocksize, first, stride):
    h, w = frame.shape[:2]
    rects = tile_rects(w, h, blocksize)
    for k, ti in enumerate(range(first, len(rects), stride)):
        x, y, tw, th = rects[ti]
        frame[y:y + th, x:x + tw] = packed[k, :tw * th].reshape((th, tw) + frame.shape[2:])
    return frame


class ShardedRenderer:
    """One rank of an N-GPU frame render.  `render_frame_dev()` leaves the complete frame on every
    rank's GPU; `render_frame_host()` additionally copies the packed 0x00RRGGBB image to pinned host
    memory (what GlomeView blits, Glome.hs:353-358)."""

    def __init__(self, scene, cam, width, height, mode, recurs, rank=0, world=1, blocksize=65, want_tcolor=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.scene, self.cam = scene, cam
        self.w, self.h, self.bs = width, height, blocksize
        self.rank, self.world = rank, world
        self.lib = L.load()
        self.opts = render_opts(mode=mode, recurs=recurs, blocksize=blocksize, tile_first=rank, tile_stride=world)
        dev = torch.device("cuda", scene.device)
        self.dev = dev
        self.tcolor = torch.zeros((height, width, 5), dtype=torch.float64, device=dev)
        self.rgb8 = torch.zeros((height, width), dtype=torch.int32, device=dev)
        self.slots = tile_slots(width, height, blocksize, world)
        self.want_tcolor = want_tcolor
        if world > 1:
            n = self.slots * blocksize * blocksize
            self.pk_rgb = torch.zeros(n, dtype=torch.int32, device=dev)
            self.ga_rgb = torch.zeros(n * world, dtype=torch.int32, device=dev)
            if want_tcolor:
                self.pk_tc = torch.zeros(n * 5, dtype=torch.float64, device=dev)
                self.ga_tc = torch.zeros(n * 5 * world, dtype=torch.float64, device=dev)
        self.host_rgb8 = torch.zeros((height, width), dtype=torch.int32).pin_memory()
        self.launches = 0
        self.last_stats = None

    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def render_frame_dev(self, want_stats=False):
        """Render this rank's tiles, gather all ranks' tiles.  Everything is enqueued on torch's current
        stream; returns the number of kernels this call launched from libglomecuda."""
        st = L.GlomeRenderStats()
        stream = self._stream()
        l0 = self.lib.glome_scene_launches(self.scene.h)
        L.check(self.lib.glome_render_dev(self.scene.h, C.byref(self.cam), self.w, self.h, C.byref(self.opts),
                                          C.c_void_p(self.tcolor.data_ptr()), C.c_void_p(self.rgb8.data_ptr()),
                                          C.byref(st) if want_stats else None, C.c_void_p(stream)))
        launches = int(self.lib.glome_scene_launches(self.scene.h) - l0)
        if want_stats:
            self.last_stats = st
        if self.world > 1:
            pairs = [(self.rgb8, self.pk_rgb, self.ga_rgb, 4)]
            if self.want_tcolor:
                pairs.append((self.tcolor, self.pk_tc, self.ga_tc, 40))
            for frame, pk, ga, eb in pairs:
                L.check(self.lib.glome_tiles_pack_dev(self.w, self.h, self.bs, self.rank, self.world, eb,
                                                      C.c_void_p(frame.data_ptr()), C.c_void_p(pk.data_ptr()),
                                                      C.c_void_p(stream)))
                self.dist.all_gather_into_tensor(ga, pk)
                L.check(self.lib.glome_tiles_unpack_all_dev(self.w, self.h, self.bs, self.world, self.rank, eb,
                                                            C.c_void_p(ga.data_ptr()), C.c_void_p(frame.data_ptr()),
                                                            C.c_void_p(stream)))
                launches += 2
        self.launches += launches
        return launches

    def render_frame_host(self, copy_on=None):
        """End-to-end frame: render + gather + device->host copy of the packed image into pinned memory.
        copy_on=None: every rank ends with the frame in host memory; copy_on=k: only rank k does (the rank that
        displays it, as GlomeView's single blit loop does), the others just wait for their part of the gather."""
        n = self.render_frame_dev()
        if copy_on is None or copy_on == self.rank:
            self.host_rgb8.copy_(self.rgb8, non_blocking=True)
        self.torch.cuda.current_stream(self.dev).synchronize()
        return self.host_rgb8, n
