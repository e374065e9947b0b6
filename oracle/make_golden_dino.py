"""Golden vectors for the DINO feature provider.  TEST INFRASTRUCTURE ONLY; build container only.

Runs the UNMODIFIED reference code -- `models/extractor.py::VitExtractor` methods on top of the vendored
`models/vision_transformer.py::vit_small` -- on seeded inputs and stores (sub-sampled) outputs in
tests/golden/dino_vits16.npz.  The reference's constructor downloads the hub model; here the instance is assembled
around the vendored ViT, loaded (strict) with the seeded random-init state_dict of nerfsos_b200's own provider, so the
test can rebuild the identical weights from the seed instead of shipping 86 MB.

    python oracle/make_golden_dino.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402  (stubs + /root/reference on sys.path)

assert ref_shim.available(), "needs /root/reference"
sys.path.insert(0, os.path.join(ref_shim.REF, "models"))     # vision_transformer.py imports its sibling `dino_utils` as a top-level module
import torch.nn as nn  # noqa: E402
from models import extractor as ref_ext  # noqa: E402
from models import vision_transformer as ref_vit  # noqa: E402

import nerfsos_b200  # noqa: E402,F401
from nerfsos_b200.models.extractor import VitExtractor  # noqa: E402

SEED = 0


def reference_extractor(state_dict):
    ext = ref_ext.VitExtractor.__new__(ref_ext.VitExtractor)      # skip torch.hub.load, keep every method
    nn.Module.__init__(ext)
    ext.model = ref_vit.vit_small(patch_size=16)
    ext.model.load_state_dict(state_dict, strict=True)
    ext.model.eval()
    ext.model_name = "dino_vits16"
    ext.hook_handlers, ext.layers_dict, ext.outputs_dict = [], {}, {}
    for key in ref_ext.VitExtractor.KEY_LIST:
        ext.layers_dict[key], ext.outputs_dict[key] = [], []
    ext._init_hooks_data()
    return ext


def inputs():
    g = torch.Generator().manual_seed(3)
    return torch.rand(2, 3, 64, 96, generator=g), torch.rand(1, 3, 64, 96, generator=g)


def main():
    ours = VitExtractor("dino_vits16", device="cpu", seed=SEED)
    ref = reference_extractor(ours.model.state_dict())
    x, xn = inputs()
    with torch.no_grad():
        a = ref.get_vit_attn_feat(x)
        b = ref.get_vit_attn_feat_noresize(xn)
        f = ref.get_vit_feature(xn)
        c = ref.get_vit_feature_attn(x)
    out = {"attn": a["attn"], "cls_": a["cls_"], "feat_sub": a["feat"][:, ::7, ::16],
           "nr_attn": b["attn"], "nr_cls_": b["cls_"], "nr_feat_sub": b["feat"][:, :, ::8],
           "vit_feature_sub": f[:, :, ::8], "vit_feature_attn": c}
    path = os.path.join(ROOT, "tests", "golden", "dino_vits16.npz")
    np.savez_compressed(path, seed=SEED, **{k: v.numpy() for k, v in out.items()})
    print(path, {k: tuple(v.shape) for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
