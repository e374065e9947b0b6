"""Where the training step's time goes outside kernels: run bench.py's train step under torch.profiler (kineto), take one
step period (render kernel to render kernel) and list GPU busy time, idle gaps and the kernels on either side of the largest gaps.
usage: python tools/profile_train_step.py [exact|fast] [--all-params]   (GPU box; prints a text report)"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "exact"
    allp = "--all-params" in sys.argv
    import nerfsos_b200  # noqa: F401
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        r = bench.train_bench(dev, 0, 0, 1, None, mode, 3, 3, all_params=allp, collect_clocks=False)
    print("ms_per_step under the profiler:", r["ms_per_step"])
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    gpu = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    gpu.sort(key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(gpu) if "k_render_tc" in e["name"] and ("1, 1" in e["name"] or "0, 1" in e["name"])]
    if len(starts) < 2:
        starts = [i for i, e in enumerate(gpu) if "k_render_tc" in e["name"]]
    a, b = starts[-2], starts[-1]
    seg = gpu[a:b]
    period = gpu[b]["ts"] - gpu[a]["ts"]
    busy, end, gaps = 0.0, seg[0]["ts"], []
    for i, e in enumerate(seg):
        if e["ts"] > end:
            gaps.append((e["ts"] - end, seg[i - 1]["name"][:70], e["name"][:70]))
        s = max(e["ts"], end)
        busy += max(0.0, e["ts"] + e["dur"] - s)
        end = max(end, e["ts"] + e["dur"])
    gaps.append((gpu[b]["ts"] - end, seg[-1]["name"][:70], "(next step's render kernel)"))
    print(f"step period {period / 1e3:.2f} ms, {len(seg)} GPU activities, busy {busy / 1e3:.2f} ms, idle {(period - busy) / 1e3:.2f} ms")
    print("idle gaps > 20 us: %d, sum %.2f ms" % (sum(g[0] > 20 for g in gaps), sum(g[0] for g in gaps if g[0] > 20) / 1e3))
    for g in sorted(gaps, reverse=True)[:25]:
        print(f"  {g[0]:8.1f} us   after {g[1]}   before {g[2]}")
    agg = {}
    for e in seg:
        k = e["name"][:70]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += e["dur"]
    print("top kernels of the step:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        print(f"  {v[0]:4d} x  {v[1] / 1e3:8.3f} ms  {k}")


if __name__ == "__main__":
    main()
