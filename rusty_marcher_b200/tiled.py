"""Row-tiled multi-GPU rendering: one process per GPU (torch.distributed), SURVEY.md 8(e).

The frame's floor(H/32) patch rows (32-row bands) are dealt round-robin to the ranks, rank k of G renders bands
G-1-k, 2G-1-k, ... (bands_of) with the same kernels as the single-GPU path.  The path has exactly one exchange step:
(1) the maximum of ONE float per rank (FrameBuffer::normalize is a global maximum, framebuffer.rs:58-69) and
(2) the normalised RGB8 rows go to rank 0.

Two implementations of that step:
  * "peer" (the product path on GPUs): the render kernel does it itself over NVLink peer memory -- rm_render_frame(), two
    kernel launches per frame and rank, no NCCL call (include/rm_b200.h, "one frame on the GPUs of one box").
    torch.distributed only carries the 64-byte CUDA IPC handles once, at set-up.
  * "collective": all_reduce(MAX) + gather through torch.distributed (gloo on CPU for the host-logic tests with a
    stand-in backend; NCCL on GPUs if CUDA IPC is not permitted on the box).
Bands are independent, so the assembled frame is bit-identical to the single-GPU frame either way.

torch is plumbing only here (device buffers, streams, rendezvous); the kernels are reached through the C ABI with raw
device pointers.
"""
import ctypes as C
import sys

import torch
import torch.distributed as dist

from . import _abi


def tile_of(n_patch_rows, rank, world):
    """Contiguous patch-row range of `rank` (tile edges coincide with the reference's patch rows)."""
    return (rank * n_patch_rows) // world, ((rank + 1) * n_patch_rows) // world


def bands_of(n_patch_rows, rank, world):
    """Interleaved row tiles: a rank renders every world-th patch row, as (begin, end, stride).  Scenes keep their work
    in one part of the frame (the cornell box's visible block sits in the top half), so contiguous tiles leave most
    ranks idle; dealing the 32-row bands round-robin balances them without any knowledge of the scene.  Rank k starts at
    band world-1-k: when the bands do not divide evenly the classes that start first get one band more, and rank 0 --
    which also receives every other rank's tiles and waits for the last of them -- should be among those with one less."""
    return world - 1 - rank, n_patch_rows, world


class _DeviceBytes:
    """A raw device allocation as something torch.as_tensor understands."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class PeerExchange:
    """This rank's view of the box's shared memory: every rank's mailbox and rank 0's two 8-bit frames, allocated with
    rm_peer_alloc, exported as CUDA IPC handles, exchanged once through torch.distributed and mapped with rm_peer_open."""

    def __init__(self, L, rank, world, frame_bytes, group=None):
        self.L, self.rank, self.world = L, rank, world
        self.owned, self.opened = [], []
        self.x = _abi.RmExchange()
        self.x.rank, self.x.world = rank, world
        box, box_h = self._alloc(_abi.RM_MAILBOX_BYTES)
        frames = [self._alloc(frame_bytes) for _ in range(2)] if rank == 0 else []
        mine = {"mailbox": box_h, "frames": [h for _p, h in frames]}
        everyone = [mine]
        if world > 1:
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
        for r in range(world):
            self.x.mailbox[r] = box if r == rank else self._open(everyone[r]["mailbox"])
        for i in range(2):
            self.x.frame8[i] = frames[i][0] if rank == 0 else self._open(everyone[0]["frames"][i])
        self.frame_ptrs = [int(self.x.frame8[0]), int(self.x.frame8[1])]

    def _alloc(self, nbytes):
        p = C.c_void_p()
        h = C.create_string_buffer(_abi.RM_IPC_HANDLE_BYTES)
        _abi.check(self.L.rm_peer_alloc(nbytes, C.byref(p), h))
        self.owned.append(p.value)
        return p.value, h.raw

    def _open(self, handle):
        p = C.c_void_p()
        _abi.check(self.L.rm_peer_open(handle, C.byref(p)))
        self.opened.append(p.value)
        return p.value

    def status(self):
        """Raises RmError(RM_ERR_PEER) if a wait on this GPU timed out (call after a synchronisation)."""
        _abi.check(self.L.rm_peer_status(C.byref(self.x)))

    def close(self):
        """Unmaps the peers' memory and frees this rank's (every rank must have stopped rendering)."""
        for p in self.opened:
            self.L.rm_peer_close(p)
        for p in self.owned:
            self.L.rm_peer_free(p)
        self.opened, self.owned = [], []


class CudaBackend:
    """Runs K0/K1/K4 through the C ABI on torch-owned device buffers, on torch's current stream."""

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
        return self.renderer.params(self._fb, self.scene, rows)

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

    def frame_params(self, rows):
        return self._params(rows)

    def render_frame(self, params, rgb, dmax, exchange, seq, normalise=True, prim=None):
        """K0 + K1 + K4 with the exchange over peer memory inside the kernels (rm_render_frame)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        pp = prim.data_ptr() if prim is not None else None
        _abi.check(self.L.rm_render_frame(self.handle, C.byref(params), rgb.data_ptr(), pp, self._ptr(dmax),
                                          C.byref(exchange.x), seq, int(normalise), stream))


class TiledRenderer:
    """Renders one frame across the ranks of `group`; rank 0 ends up with the whole RGB8 frame.

    exchange: "peer" (kernels exchange over NVLink peer memory, needs a backend with render_frame), "collective"
    (torch.distributed all_reduce + gather), or "auto": peer when the backend can, falling back to collective -- loudly --
    only if CUDA IPC cannot be set up on this box."""

    def __init__(self, backend, width, height, device, group=None, exchange="auto"):
        self.backend = backend
        self.width, self.height = width, height
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_patch_rows = height // 32
        self.rows = bands_of(self.n_patch_rows, self.rank, self.world) if self.world > 1 else (0, self.n_patch_rows)
        self.seq = 0
        # full-frame float buffer (only this rank's rows are written) and the max scalar
        self.rgb = torch.zeros((height, width, 3), dtype=torch.float32, device=device)
        self.dmax = torch.zeros(1, dtype=torch.float32, device=device)
        self.peer = None
        self.exchange = exchange
        if exchange not in ("auto", "peer", "collective"):
            raise ValueError("exchange must be 'auto', 'peer' or 'collective'")
        if exchange != "collective" and hasattr(backend, "render_frame"):
            self.exchange = self._setup_peer(strict=(exchange == "peer"))
        elif exchange == "peer":
            raise ValueError("this backend has no render_frame(): the peer exchange needs the CUDA backend")
        else:
            self.exchange = "collective"
        if self.exchange == "collective":
            self._setup_collective()

    # ---- peer exchange ------------------------------------------------------------------------------------------------
    def _setup_peer(self, strict):
        ok, err = 1, ""
        try:
            self.peer = PeerExchange(self.backend.L, self.rank, self.world, self.height * self.width * 3, self.group)
        except Exception as e:                                   # noqa: BLE001 -- decided collectively below
            if self.world == 1 or strict:
                raise
            ok, err = 0, str(e)
        if self.world > 1:
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ok = int(flag.item())
        if not ok:
            if self.peer is not None:
                self.peer.close()
                self.peer = None
            sys.stderr.write("rusty_marcher_b200.tiled: CUDA IPC peer mapping failed on a rank (%s); "
                             "falling back to the torch.distributed exchange\n" % (err or "see the other ranks"))
            return "collective"
        self.params = self.backend.frame_params(self.rows)
        n = self.height * self.width * 3
        self.frames8 = None
        if self.rank == 0:
            self.frames8 = [torch.as_tensor(_DeviceBytes(p, n), device=self.device).view(self.height, self.width, 3)
                            for p in self.peer.frame_ptrs]
        return "peer"

    # ---- torch.distributed exchange -----------------------------------------------------------------------------------
    def _setup_collective(self):
        height, width, device = self.height, self.width, self.device
        self.rgb8 = torch.zeros((height, width, 3), dtype=torch.uint8, device=device)
        # [patch row, 32, W, 3] view of the rendered part: rank r owns [firsts[r]::world]
        self.bands8 = self.rgb8[:self.n_patch_rows * 32].view(self.n_patch_rows, 32, width, 3)
        self.firsts = [bands_of(self.n_patch_rows, r, self.world)[0] for r in range(self.world)]
        self.counts = [len(range(f, self.n_patch_rows, self.world)) for f in self.firsts]
        # equal-size gather slots: ranks with one band less pad
        self.slot = torch.zeros((max(self.counts + [1]), 32, width, 3), dtype=torch.uint8, device=device)
        self.gathered = ([torch.zeros_like(self.slot) for _ in range(self.world)] if self.rank == 0 and self.world > 1 else None)

    def render(self):
        """One frame, asynchronous on the current stream.  Returns the device RGB8 frame on rank 0 (None elsewhere);
        with the peer exchange consecutive frames alternate between two buffers."""
        self.seq += 1
        if self.exchange == "peer":
            self.backend.render_frame(self.params, self.rgb, self.dmax, self.peer, self.seq)
            return self.frames8[self.seq & 1] if self.rank == 0 else None
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
            self.slot[:n].copy_(self.bands8[self.firsts[self.rank]::self.world])
        dist.gather(self.slot, self.gathered, dst=0, group=self.group)
        if self.rank != 0:
            return None
        for r in range(1, self.world):
            if self.counts[r]:
                self.bands8[self.firsts[r]::self.world].copy_(self.gathered[r][:self.counts[r]])
        return self.rgb8

    def set_camera(self, camera):
        """Camera of the following frames (the scene stays resident on the device; main.rs:124-171 moves the camera only)."""
        if self.exchange == "peer":
            self.params.camera[:] = [float(c) for c in camera]
        else:
            self.backend.scene.camera = type(self.backend.scene.camera)(*[float(c) for c in camera])

    def launches_per_frame(self):
        """Kernels of this repo launched per frame on this rank: K0 and K1 (K4 fused) on the peer path, K0, K1, K4 otherwise."""
        return 2 if self.exchange == "peer" else 3

    def close(self):
        if self.peer is not None:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier(group=self.group)
            self.frames8 = None
            self.peer.close()
            self.peer = None
