"""Row-tiled multi-GPU rendering: one process per GPU (torch.distributed), SURVEY.md 8(e).

The frame's floor(H/32) patch rows are split into contiguous tiles, rank k of G renders patch rows
[k*P//G, (k+1)*P//G) with the same kernels as the single-GPU path, then the only exchange step of
the path runs: (1) a max all-reduce of ONE float (FrameBuffer::normalize is a global maximum,
framebuffer.rs:58-69) and (2) a gather of the normalised RGB8 rows to rank 0 over NVLink (NCCL).
Tiles are independent, so the assembled frame is bit-identical to the single-GPU frame.

torch is plumbing only here (device buffers, streams, NCCL); the kernels are reached through the
C ABI with raw device pointers.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _abi


def tile_of(n_patch_rows, rank, world):
    """Contiguous patch-row range of `rank` (tile edges coincide with the reference's patch rows)."""
    return (rank * n_patch_rows) // world, ((rank + 1) * n_patch_rows) // world


def bands_of(n_patch_rows, rank, world):
    """Interleaved row tiles: rank k renders patch rows k, k + world, ... as (begin, end, stride).  Scenes keep their
    work in one part of the frame (the cornell box's visible block sits in the top half), so contiguous tiles leave
    most ranks idle; dealing the 32-row bands round-robin balances them without any knowledge of the scene."""
    return rank, n_patch_rows, world


class CudaBackend:
    """Runs K1/K4 through the C ABI on torch-owned device buffers, on torch's current stream."""

    def __init__(self, scene, renderer, width, height, device):
        self.scene, self.renderer = scene, renderer
        self.width, self.height = width, height
        self.device = device
        self.L = _abi.load()
        _abi.init(device.index if device.index is not None else 0)
        self.handle = scene.device_handle()
        from .framebuffer import FrameBuffer
        self._fb = FrameBuffer.__new__(FrameBuffer)
        self._fb.width, self._fb.height = width, height

    def _params(self, rows):
        """rows = (begin, end) or (begin, end, stride) in patch rows."""
        p = self.renderer.params(self._fb, self.scene, rows[:2])
        p.patch_row_stride = rows[2] if len(rows) > 2 else 1
        return p

    @staticmethod
    def _ptr(t):
        return t if isinstance(t, int) else t.data_ptr()

    def render_rows(self, rows, rgb, dmax, prim=None, rgb8=None):
        """K0 + K1.  With rgb8 (tensor or raw device pointer, possibly peer-mapped) the render kernel also zeroes the
        8-bit frame where it can (rm_render_device_rgb8), and tonemap_rows(..., busy=True) converts only the rest."""
        p = self._params(rows)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        pp = prim.data_ptr() if prim is not None else None
        if rgb8 is None:
            _abi.check(self.L.rm_render_device(self.handle, C.byref(p), rgb.data_ptr(), pp, self._ptr(dmax), stream))
        else:
            _abi.check(self.L.rm_render_device_rgb8(self.handle, C.byref(p), rgb.data_ptr(), pp, self._ptr(dmax),
                                                    self._ptr(rgb8), stream))

    def tonemap_rows(self, rows, rgb, dmax, rgb8, normalise=True, busy=False):
        p = self._params(rows)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if busy:
            _abi.check(self.L.rm_tonemap_device_busy(self.handle, C.byref(p), rgb.data_ptr(), self._ptr(dmax), int(normalise),
                                                     self._ptr(rgb8), stream))
        else:
            _abi.check(self.L.rm_tonemap_device(C.byref(p), rgb.data_ptr(), self._ptr(dmax), int(normalise),
                                                self._ptr(rgb8), stream))


class TiledRenderer:
    """Renders one frame across the ranks of `group`; rank 0 ends up with the whole RGB8 frame.

    Exchange step over NCCL: a one-float max all-reduce, then a gather of every rank's bands (packed) to rank 0,
    which scatters them back to their rows."""

    def __init__(self, backend, width, height, device, group=None, keep_float=False):
        self.backend = backend
        self.width, self.height = width, height
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_patch_rows = height // 32
        self.rows = bands_of(self.n_patch_rows, self.rank, self.world) if self.world > 1 else (0, self.n_patch_rows)
        # full-frame float buffer (only this rank's rows are written), the max scalar, the RGB8 frame
        self.rgb = torch.zeros((height, width, 3), dtype=torch.float32, device=device)
        self.dmax = torch.zeros(1, dtype=torch.float32, device=device)
        self.rgb8 = torch.zeros((height, width, 3), dtype=torch.uint8, device=device)
        # [patch row, 32, W, 3] view of the rendered part: rank r owns [r::world]
        self.bands8 = self.rgb8[:self.n_patch_rows * 32].view(self.n_patch_rows, 32, width, 3)
        self.counts = [len(range(r, self.n_patch_rows, self.world)) for r in range(self.world)]
        # equal-size gather slots: ranks with one band less pad
        self.slot = torch.zeros((max(self.counts + [1]), 32, width, 3), dtype=torch.uint8, device=device)
        self.gathered = ([torch.zeros_like(self.slot) for _ in range(self.world)] if self.rank == 0 and self.world > 1 else None)

    def render(self):
        """One frame.  Returns the device RGB8 frame on rank 0 (None elsewhere)."""
        self.dmax.zero_()
        if self.world == 1:
            self.backend.render_rows(self.rows, self.rgb, self.dmax, rgb8=self.rgb8)
            self.backend.tonemap_rows(self.rows, self.rgb, self.dmax, self.rgb8, busy=True)
            return self.rgb8
        self.backend.render_rows(self.rows, self.rgb, self.dmax)
        dist.all_reduce(self.dmax, op=dist.ReduceOp.MAX, group=self.group)
        self.backend.tonemap_rows(self.rows, self.rgb, self.dmax, self.rgb8)
        n = self.counts[self.rank]
        if n:
            self.slot[:n].copy_(self.bands8[self.rank::self.world])
        dist.gather(self.slot, self.gathered, dst=0, group=self.group)
        if self.rank != 0:
            return None
        for r in range(1, self.world):
            if self.counts[r]:
                self.bands8[r::self.world].copy_(self.gathered[r][:self.counts[r]])
        return self.rgb8
