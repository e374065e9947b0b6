"""How much of the training forward is the cost of saving activations: the same 32768 rays with the training kwargs, rendered
under no_grad (kernel A MODE 0) and with the semantic heads requiring grad (MODE 1: h_last / s_hid / gamma / raw written).
usage: python tools/time_train_fwd.py [exact|fast]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200.models.nerf_net import NeRFNet
    dev = torch.device("cuda", 0)
    n = 8 * 64 * 64
    net = NeRFNet(N_samples=bench.N_SAMPLES, N_importance=bench.N_IMPORTANCE, use_semantics=True, sem_with_coord=True, sem_dim=2,
                  perturb=1.0, raw_noise_std=1.0, mode=mode)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights().items()}, strict=True)
    net = net.to(dev).train()
    for name, p in net.named_parameters():
        p.requires_grad_("semantic_linear" in name)
    rays = torch.from_numpy(bench.llff_rays(n, 200)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(grad):
        ts = []
        for i in range(6):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.set_grad_enabled(grad):
                out = net((rays[0], rays[1]), (0.0, 1.0))
            e1.record()
            torch.cuda.synchronize()
            del out
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    print(f"mode={mode} rays={n}: no_grad {run(False):.2f} ms, saving activations {run(True):.2f} ms")


if __name__ == "__main__":
    main()
