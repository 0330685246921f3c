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
        return self.renderer.params(self._fb, self.scene, rows)

    def render_rows(self, rows, rgb, dmax, prim=None):
        p = self._params(rows)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _abi.check(self.L.rm_render_device(self.handle, C.byref(p), rgb.data_ptr(),
                                           prim.data_ptr() if prim is not None else None, dmax.data_ptr(), stream))

    def tonemap_rows(self, rows, rgb, dmax, rgb8, normalise=True):
        p = self._params(rows)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _abi.check(self.L.rm_tonemap_device(C.byref(p), rgb.data_ptr(), dmax.data_ptr(), int(normalise),
                                            rgb8.data_ptr(), stream))


class TiledRenderer:
    """Renders one frame across the ranks of `group`; rank 0 ends up with the whole RGB8 frame."""

    def __init__(self, backend, width, height, device, group=None, keep_float=False):
        self.backend = backend
        self.width, self.height = width, height
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_patch_rows = height // 32
        self.rows = tile_of(self.n_patch_rows, self.rank, self.world)
        # full-frame float buffer (only this rank's rows are written), the max scalar, this rank's RGB8 rows
        self.rgb = torch.zeros((height, width, 3), dtype=torch.float32, device=device)
        self.dmax = torch.zeros(1, dtype=torch.float32, device=device)
        self.rgb8 = torch.zeros((height, width, 3), dtype=torch.uint8, device=device)
        self.tiles = [tile_of(self.n_patch_rows, r, self.world) for r in range(self.world)]
        max_rows = max(b - a for a, b in self.tiles) * 32
        # equal-size gather slots: ranks with one patch row less pad
        self.slot = torch.zeros((max_rows, width, 3), dtype=torch.uint8, device=device)
        self.gathered = ([torch.zeros_like(self.slot) for _ in range(self.world)] if self.rank == 0 and self.world > 1 else None)

    def render(self):
        """One frame.  Returns the device RGB8 frame on rank 0 (None elsewhere)."""
        a, b = self.rows
        self.dmax.zero_()
        self.backend.render_rows((a, b), self.rgb, self.dmax)
        if self.world > 1:
            dist.all_reduce(self.dmax, op=dist.ReduceOp.MAX, group=self.group)
        self.backend.tonemap_rows((a, b), self.rgb, self.dmax, self.rgb8)
        if self.world == 1:
            return self.rgb8
        n = (b - a) * 32
        self.slot[:n].copy_(self.rgb8[a * 32:b * 32])
        dist.gather(self.slot, self.gathered, dst=0, group=self.group)
        if self.rank != 0:
            return None
        for r, (ra, rb) in enumerate(self.tiles):
            if r != 0:
                self.rgb8[ra * 32:rb * 32].copy_(self.gathered[r][:(rb - ra) * 32])
        return self.rgb8
