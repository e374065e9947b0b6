"""Host-side cost of one end-to-end eval call (pinned rays in, maps out): cProfile over 300 calls, top entries by own time.
usage: python tools/profile_eval_host.py [exact|fast]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200.models.nerf_net import NeRFNet
    dev = torch.device("cuda", 0)
    net = NeRFNet(N_samples=bench.N_SAMPLES, N_importance=bench.N_IMPORTANCE, use_semantics=True, sem_with_coord=True, sem_dim=2, mode=mode)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights().items()}, strict=True)
    net = net.to(dev).eval()
    rays_host = torch.from_numpy(bench.llff_rays(4096, 0)).pin_memory()
    maps_host = torch.empty(4096, 17).pin_memory()

    def step():
        with torch.no_grad():
            r = rays_host.to(dev, non_blocking=True)
            out = net(r, (bench.NEAR, bench.FAR), retraw=False, retmaps=True)
            maps_host.copy_(out["maps"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(20):
        step()
    import time
    t0 = time.perf_counter()
    for _ in range(300):
        step()
    dt = (time.perf_counter() - t0) / 300
    print(f"wall per call {dt * 1e3:.3f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        step()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__" and "timeline" not in sys.argv:
    main()


def timeline():
    """GPU-side timeline of single e2e calls (kineto): H2D copy, launch gap, kernel, D2H copy."""
    import json
    import tempfile
    from torch.profiler import ProfilerActivity, profile
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200.models.nerf_net import NeRFNet
    dev = torch.device("cuda", 0)
    net = NeRFNet(N_samples=bench.N_SAMPLES, N_importance=bench.N_IMPORTANCE, use_semantics=True, sem_with_coord=True, sem_dim=2, mode="exact")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights().items()}, strict=True)
    net = net.to(dev).eval()
    rays_host = torch.from_numpy(bench.llff_rays(4096, 0)).pin_memory()
    maps_host = torch.empty(4096, 17).pin_memory()

    def step():
        with torch.no_grad():
            r = rays_host.to(dev, non_blocking=True)
            out = net(r, (bench.NEAR, bench.FAR), retraw=False, retmaps=True)
            maps_host.copy_(out["maps"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(10):
        step()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(6):
            step()
    path = os.path.join(tempfile.mkdtemp(), "t.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X"]
    gpu = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
    sync = sorted([e for e in ev if e.get("name", "").startswith("cudaStreamSynchronize")], key=lambda e: e["ts"])
    h2d = [e for e in gpu if "HtoD" in e["name"]]
    for i in range(1, len(h2d) - 1):
        t0 = h2d[i]["ts"]
        seg = [e for e in gpu if t0 <= e["ts"] < h2d[i + 1]["ts"]]
        s = [x for x in sync if x["ts"] >= t0]
        end = s[0]["ts"] + s[0]["dur"] if s else 0
        print(f"call {i}: " + "  ".join(f"{e['name'][:18]}@{e['ts'] - t0:.0f}+{e['dur']:.0f}us" for e in seg) + f"  | sync returns @{end - t0:.0f}us; next call's H2D @{h2d[i + 1]['ts'] - t0:.0f}us")


if __name__ == "__main__" and "timeline" in sys.argv:
    timeline()
