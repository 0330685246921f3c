"""Device-to-pinned-host copy time vs size, back to back and after idle gaps (what bounds the host delivery of a frame)."""
import time

import torch

dev = torch.device("cuda:0")
for mb in (1.2, 2.4, 7.2, 14.4, 28.8, 98.8):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for gap_ms in (0.0, 1.0):
        ts, es = [], []
        for i in range(25):
            if gap_ms:
                time.sleep(gap_ms * 1e-3)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record()
            h.copy_(d, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            es.append(a.elapsed_time(b))
        ts.sort()
        es.sort()
        print("%6.1f MB gap %.0f ms: wall median %.3f ms (%.1f GB/s), events median %.3f ms (%.1f GB/s)" % (
            mb, gap_ms, 1e3 * ts[12], mb / ts[12] / 1e3, es[12], mb / es[12]))
