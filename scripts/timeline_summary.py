#!/usr/bin/env python
"""Summarise a B200PT_TIMELINE dump (frame,lane,stage,start_ms,end_ms: CUDA-event pairs around every launch of a REAL run,
frames in flight not serialised).  Prints, over the steady-state window: wall span per frame, the time at least one kernel
of ours was running, mean number of kernels in flight, per-stage sums and the per-frame critical path of a lane.

    python scripts/timeline_summary.py gpurun_out/tl.csv.<pid> [skip_frames]
"""
import sys
from collections import defaultdict

import numpy as np


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rows = []
    for line in open(path).read().splitlines()[1:]:
        f = line.split(",")
        if len(f) == 5:
            rows.append((int(f[0]), int(f[1]), f[2], float(f[3]), float(f[4])))
    if not rows:
        print("empty timeline")
        return
    frames = sorted({r[0] for r in rows})
    # (a 'frame' is one launch chain = one batch of frames; the first ones are warm-up)
    keep = set(frames[skip:-1]) if len(frames) > skip + 3 else (set(frames[skip:]) if len(frames) > skip else set(frames))
    rows = [r for r in rows if r[0] in keep]
    t0, t1 = min(r[3] for r in rows), max(r[4] for r in rows)
    n_frames = len(keep)
    ev = []
    for r in rows:
        ev.append((r[3], 1))
        ev.append((r[4], -1))
    ev.sort()
    busy, depth_time, cur, last = 0.0, 0.0, 0, ev[0][0]
    for t, d in ev:
        if cur > 0:
            busy += t - last
        depth_time += cur * (t - last)
        cur += d
        last = t
    span = t1 - t0
    print(f"{path}: {n_frames} launch chains (batches), {len(rows)} launches, span {span:.2f} ms = {span / n_frames:.3f} ms per chain")
    print(f"  some kernel of ours running: {100 * busy / span:.1f} % of the span; mean kernels in flight {depth_time / span:.2f}")
    per = defaultdict(lambda: [0.0, 0])
    for r in rows:
        per[r[2]][0] += r[4] - r[3]
        per[r[2]][1] += 1
    tot = sum(v[0] for v in per.values())
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:10s} {v[0] / n_frames:8.3f} ms/frame summed ({100 * v[0] / tot:5.1f} %), {v[1] / n_frames:6.1f} launches/frame, {1e3 * v[0] / v[1]:8.1f} us/launch")
    lat = defaultdict(lambda: [1e30, -1e30])
    for r in rows:
        lat[r[0]][0] = min(lat[r[0]][0], r[3])
        lat[r[0]][1] = max(lat[r[0]][1], r[4])
    l = np.array([b - a for a, b in lat.values()])
    print(f"  frame latency (first launch -> accumulate): mean {l.mean():.2f} ms, max {l.max():.2f} ms; sum of a frame's kernel times {tot / n_frames:.2f} ms")


if __name__ == "__main__":
    main()
