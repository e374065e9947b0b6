#!/usr/bin/env python
"""Kernel A time vs ray count (296 rays = one pair per CTA on 148 SMs): separates the per-pair cost from launch-level
overheads.  python tools/scan_rays.py [exact|fast]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nerfsos_b200  # noqa
from tools_common import make_net
import bench
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
net = make_net(mode)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = []
for n in [296 * k for k in (1, 2, 3, 4, 6, 8, 10, 12, 13, 14)] + [4096, 296 * 28]:
    rays = torch.from_numpy(bench.llff_rays(n, 100)).cuda()
    ts = {}
    for warm in (True, False):
        ms = []
        with torch.no_grad():
            for i in range(8):
                if not warm:
                    flush.fill_(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); net(rays, (1.2, 12.0), retraw=False); e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
        ts[warm] = float(np.median(ms[3:]))
    res.append((n, ts[True], ts[False]))
    print(f"rays {n:6d} pairs/CTA {n / 296:6.2f}: {ts[True]*1e3:8.1f} us (L2 warm) {ts[False]*1e3:8.1f} us (L2 flushed)  -> {ts[False]*1e3/(n/296):7.1f} us per pair-iteration")
a = np.array(res)
k = a[:10, 0] / 296
for col, name in ((1, "warm"), (2, "flushed")):
    slope, icpt = np.polyfit(k, a[:10, col] * 1e3, 1)
    print(f"{name}: {slope:.1f} us per pair-iteration + {icpt:.1f} us fixed")
