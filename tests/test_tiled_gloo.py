"""Multi-rank host logic on CPU: world_size 2 and 3 over gloo.  The row-tile
partition, the one-float max all-reduce and the RGB8 gather of rusty_marcher_b200.tiled are run
with a stand-in backend (the host emulation of the kernel code) and the frame assembled on rank 0
must be byte-identical to the single-rank frame."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rusty_marcher_b200 import tiled, workloads

W, H = 128, 160     # 5 patch rows: uneven tiles for 2 and 3 ranks


class EmuBackend:
    """Stand-in for tiled.CudaBackend: same interface, kernels emulated on the host (tests only)."""

    def __init__(self, scene):
        self.scene = scene

    @staticmethod
    def _bands(rows):
        stride = rows[2] if len(rows) > 2 else 1
        return [(p, p + 1) for p in range(rows[0], rows[1], stride)]

    def render_rows(self, rows, rgb, dmax, prim=None, rgb8=None):
        from tests.emu import emu
        for pr in self._bands(rows):
            r = emu.render(self.scene, W, H, "f32", patch_rows=pr, threads=2)
            a, b = pr[0] * 32, pr[1] * 32
            rgb[a:b] = torch.from_numpy(r["rgb"][a:b])
            dmax[0] = max(float(dmax[0]), float(r["rgb"][a:b].max()))

    def tonemap_rows(self, rows, rgb, dmax, rgb8, normalise=True, busy=False):
        m = np.float32(dmax[0].item())
        inv = np.float32(1) / m if (normalise and m > 0) else np.float32(1)
        for pr in self._bands(rows):
            a, b = pr[0] * 32, pr[1] * 32
            x = rgb[a:b].numpy()
            rgb8[a:b] = torch.from_numpy((np.float32(255) * np.clip(x * inv, 0, 1)).astype(np.uint8))


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = workloads.scene("demo")
        tr = tiled.TiledRenderer(EmuBackend(scene), W, H, torch.device("cpu"))
        frame = tr.render()
        frame2 = tr.render()                                    # a second frame reuses every buffer
        if rank == 0:
            assert torch.equal(frame, frame2)
            np.save(out_path, frame.numpy())
        else:
            assert frame is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_tile_partition_covers_all_patch_rows():
    for p in (1, 5, 67, 135):
        for g in (1, 2, 3, 4, 8):
            tiles = [tiled.tile_of(p, r, g) for r in range(g)]
            assert tiles[0][0] == 0 and tiles[-1][1] == p
            assert all(tiles[i][1] == tiles[i + 1][0] for i in range(g - 1))
            sizes = [b - a for a, b in tiles]
            assert max(sizes) - min(sizes) <= 1
    assert tiled.tile_of(67, 3, 8) == (25, 33)
    for p in (1, 5, 67, 135):
        for g in (1, 2, 3, 4, 8):
            seen = sorted(r for k in range(g) for r in range(*tiled.bands_of(p, k, g)))
            assert seen == list(range(p))


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_tiled_frame_equals_single_rank_frame(tmp_path, world):
    single = tiled.TiledRenderer(EmuBackend(workloads.scene("demo")), W, H, torch.device("cpu")).render().numpy().copy()
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    multi = np.load(out)
    assert multi.shape == (H, W, 3) and multi.dtype == np.uint8
    assert np.array_equal(multi, single)
    assert multi[:H // 32 * 32].max() == 255
