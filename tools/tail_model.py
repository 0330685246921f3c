"""Load-balance model of the hierarchy kernel's schedule (no GPU needed).

The host emulation gives the walk cost of every pixel (node visits + 0.7 x primitive tests) for stage A and for the
shading stage; this script replays the kernel's work distribution on them -- 2368 warps drawing 32x2 strips from one
counter, a lane walking its two pixels one after the other, full rounds of 32 queued hits shaded as soon as a warp has
them, leftovers pooled per CTA -- with a warp's step costing the MAXIMUM over its lanes (SIMT), and reports the makespan
against the perfectly balanced time.  Alternatives are replayed on the same costs: shading rounds handed to whichever warp
is free (a global queue), and rounds formed from cost-sorted hits.

    python tools/tail_model.py [workload] [width height]
"""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload_of  # noqa: E402
from rusty_marcher_b200 import workloads  # noqa: E402
from tests.emu import emu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "stress_4k_bvh"
scene_name, w, h, depth, kw, _accel = workload_of(name)
if len(sys.argv) > 3:
    w, h = int(sys.argv[2]), int(sys.argv[3])
scene = workloads.scene(scene_name, **kw)
r = emu.render_costs(scene, w, h, max_depth=depth)
cost, hit = r["cost"], r["prim_id"] >= 0
rows = (h // 32) * 32
N_WARPS, WARPS_PER_CTA = 148 * 2 * 8, 8
STRIP_FIXED = 12.0          # zero-fill, counter, queue bookkeeping of a strip, in node-visit units
ROUND_FIXED = 20.0          # per shading round: entry decode, material loads, stores

# strips in the order the kernel draws them: tile by tile (row-major), 16 strips of 32x2 per tile
strips = []
for ty in range(rows // 32):
    for tx in range(w // 32):
        for s in range(16):
            y0, x0 = ty * 32 + 2 * s, tx * 32
            a = cost[y0:y0 + 2, x0:x0 + 32, 0]
            lane = a.reshape(2, 16, 2).sum(axis=2)                      # a lane owns two adjacent pixels of one row
            hb = cost[y0:y0 + 2, x0:x0 + 32, 1][hit[y0:y0 + 2, x0:x0 + 32]]
            strips.append((STRIP_FIXED + float(lane.max()), float(lane.sum()), hb))


def replay(global_rounds=False, sort_rounds=False):
    """Returns (makespan, sum of busy warp time, useful lane work)."""
    free = [(0.0, k) for k in range(N_WARPS)]
    heapq.heapify(free)
    queues = [[] for _ in range(N_WARPS)]
    pool = []                                                           # global_rounds: every hit of the frame
    busy = 0.0
    lane_work = 0.0
    for (ta, wa, hb) in strips:
        t, k = heapq.heappop(free)
        dt = ta
        lane_work += wa
        if global_rounds:
            pool.extend(hb.tolist())
        else:
            q = queues[k]
            q.extend(hb.tolist())
            while len(q) >= 32:
                take = q[-32:]
                del q[-32:]
                dt += ROUND_FIXED + max(take)
                lane_work += sum(take)
        busy += dt
        heapq.heappush(free, (t + dt, k))
    if global_rounds:
        if sort_rounds:
            pool.sort(reverse=True)
        for i in range(0, len(pool), 32):
            take = pool[i:i + 32]
            t, k = heapq.heappop(free)
            dt = ROUND_FIXED + max(take)
            lane_work += sum(take)
            busy += dt
            heapq.heappush(free, (t + dt, k))
    else:
        # leftovers pooled per CTA once its warps have run out of strips: one more round per 32 pooled hits
        ends = dict((k, t) for t, k in free)
        free2 = []
        for c in range(N_WARPS // WARPS_PER_CTA):
            ws = range(c * WARPS_PER_CTA, (c + 1) * WARPS_PER_CTA)
            t0 = max(ends[k] for k in ws)                               # the pooling barrier
            left = [x for k in ws for x in queues[k]]
            for i, k in enumerate(ws):
                take = left[32 * i:32 * i + 32]
                dt = (ROUND_FIXED + max(take)) if take else 0.0
                lane_work += sum(take)
                busy += dt + (t0 - ends[k])                             # waiting at the barrier keeps the warp resident
                free2.append((t0 + dt, k))
        free = free2
    return max(t for t, _ in free), busy, lane_work


print("%s at %dx%d: %d strips, %d hits, mean stage-A cost %.1f / px, mean shading cost %.1f / hit (max %.0f)" % (
    name, w, h, len(strips), int(hit[:rows].sum()), float(cost[:rows, :, 0].mean()), float(cost[..., 1][hit].mean()), float(cost[..., 1].max())))
for label, kwargs in (("kernel's schedule (per-warp queues, CTA pooling)", {}),
                      ("shading rounds from one global queue", dict(global_rounds=True)),
                      ("global queue, rounds formed from cost-sorted hits", dict(global_rounds=True, sort_rounds=True))):
    span, busy, lane_work = replay(**kwargs)
    print("%-52s makespan %9.0f  warp slots busy %5.1f %%  lanes doing useful work while busy %5.1f %%" % (
        label, span, 100.0 * busy / (span * N_WARPS), 100.0 * lane_work / (32.0 * busy)))
