import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nerfsos_b200
from nerfsos_b200.models.nerf_net import NeRFNet
from conftest import load_golden


def make_net(mode, dev="cuda:0"):
    sd = load_golden("flower_weights")["sd"]
    net = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, mode=mode)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return net.to(dev).eval()
