#!/usr/bin/env python
"""Per-stage timeline of CTA 0 of kernel A (NSOS_TRACE=1): where the cycles of a tile go."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["NSOS_TRACE"] = "1"
import nerfsos_b200
from tools_common import make_net, load_golden   # noqa
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
train = "train" in sys.argv[2:]              # training kwargs (perturb = 1, raw_noise_std = 1, retraw) under no_grad
net = make_net(mode)
if train:
    net.train()
    net.render_kwargs_train.update(perturb=1.0, raw_noise_std=1.0)
g = load_golden("flower_eval_256")
rays = torch.from_numpy(np.tile(g["rays"], (1, 16, 1))).cuda()
with torch.no_grad():
    for _ in range(3):
        net(rays, (1.2, 12.0))
torch.cuda.synchronize()
NS = 12
tr = net._ws[:16 * 16 * NS * 8].view(torch.int64).reshape(16, 16, NS).cpu().numpy()
print(f"mode={mode}: stamps of CTA 0 (cycles). per stage: acc_wait = worker waits for MMA; epi = epilogue (thread 0);")
print("mma_lead = a_ready seen by MMA warp -> accumulator ready (MMA execution incl. issue); issue = MMA warp issue time")
for tile in range(8):
    t = tr[tile]
    if t[0, 1] == 0: break
    nst = int((t[:10, 1] > 0).sum())
    base = t[0, 0]
    rows = []
    for st in range(nst):
        w0, acc, epi, ar, iss = t[st, 0], t[st, 1], t[st, 2], t[st, 3], t[st, 4]
        rows.append((st, acc - w0, epi - acc, acc - ar, iss - ar))
    tot = t[nst - 1, 2] - t[0, 0]
    print(f"tile {tile}: {nst} stages, total {tot} cycles;  sum acc_wait {sum(r[1] for r in rows)}  sum epi {sum(r[2] for r in rows)}  sum mma_exec {sum(r[3] for r in rows)}")
    if tile in (1, 2, 3, 4):
        for r in rows:
            st = r[0]
            acc = t[st, 1]          # accumulator of stage st ready = its epilogue starts; slabs are then handed to stage st+1
            seen = [int(t[st + 1, 5 + j] - acc) if st + 1 < nst and t[st + 1, 5 + j] else None for j in range(4)]
            gave = [int(t[st, 9 + j] - acc) if t[st, 9 + j] else None for j in range(3)]
            print(f"    stage {st:2d}: acc_wait {r[1]:6d}  epi {r[2]:6d}  mma_exec {r[3]:6d}  issue {r[4]:6d}   slab handed over at {gave}, seen by the next stage's MMA lane at {seen} (cycles after this epilogue started)")
    print(f"    setup (z, point, gamma -> smem) {t[15, 1] - t[15, 0]};  first acc wait starts {t[0, 0] - t[15, 1]} after a_ready;  tail (head combine, raw) {t[15, 2] - t[nst - 1, 2]}")
    if tile + 1 < 16 and tr[tile + 1][15, 0] > 0:
        print(f"    end of tile -> next tile's setup start (composite / resample / pair setup): {tr[tile + 1][15, 0] - t[15, 2]}")
    if tile + 1 < 16 and tr[tile + 1][0, 0] > 0:
        print(f"    gap to next tile's first wait (setup/composite): {tr[tile + 1][0, 0] - t[nst - 1, 2]}")


g = tr[15].reshape(-1)
print(f"pair setup (2nd pair of CTA 0): rays+dir encoding {g[41]-g[40]}  dir bias {g[42]-g[41]}")
for ps, name in ((0, "coarse"), (1, "fine")):
    d = g[ps * 16: ps * 16 + 16]
    if d[8] == 0: continue
    print(f"post-{name} phase (thread 0): composite {d[9]-d[8]}  total {d[10]-d[8]};  composite split: alpha {d[11]-d[8]}  scan {d[12]-d[11]}  products {d[13]-d[12]}  sums {d[14]-d[13]}  write {d[9]-d[14]}")
r = g[56:60]
if r[0] > 0:
    print(f"resample split: bins+cdf {r[1]-r[0]}  inverse cdf {r[2]-r[1]}  z_std {r[3]-r[2]}  merge+write {g[10]-r[3]}")
