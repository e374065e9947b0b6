"""DINO feature provider against golden vectors produced by the UNMODIFIED reference
(`models/extractor.py` methods on the vendored `models/vision_transformer.py`; oracle/make_golden_dino.py).  The weights
are rebuilt from the seed stored in the fixture, so no checkpoint ships."""
import numpy as np
import pytest
import torch

import nerfsos_b200  # noqa: F401
from conftest import load_golden
from nerfsos_b200.models.extractor import VitExtractor


def _inputs():
    g = torch.Generator().manual_seed(3)
    return torch.rand(2, 3, 64, 96, generator=g), torch.rand(1, 3, 64, 96, generator=g)


def _check(ext, dev, tol):
    gold = load_golden("dino_vits16")
    x, xn = (t.to(dev) for t in _inputs())
    a = ext.get_vit_attn_feat(x)
    assert a["attn"].shape == (2, 1, 196) and a["cls_"].shape == (2, 384) and a["feat"].shape == (2, 196, 384)
    b = ext.get_vit_attn_feat_noresize(xn)
    got = {"attn": a["attn"], "cls_": a["cls_"], "feat_sub": a["feat"][:, ::7, ::16],
           "nr_attn": b["attn"], "nr_cls_": b["cls_"], "nr_feat_sub": b["feat"][:, :, ::8],
           "vit_feature_sub": ext.get_vit_feature(xn)[:, :, ::8], "vit_feature_attn": ext.get_vit_feature_attn(x)}
    for k, v in got.items():
        assert not v.requires_grad
        np.testing.assert_allclose(v.cpu().numpy(), gold[k], rtol=tol, atol=tol, err_msg=k)
    # attention rows are probabilities over CLS + patches: the patch columns sum to < 1
    assert float(a["attn"].sum(-1).max()) < 1.0


def test_extractor_matches_reference_cpu():
    gold = load_golden("dino_vits16")
    ext = VitExtractor("dino_vits16", device="cpu", seed=int(gold["seed"]))
    assert not ext.pretrained and ext.get_patch_size() == 16 and ext.get_head_num() == 6 and ext.get_embedding_dim() == 384
    assert ext.get_patch_num((1, 3, 224, 224)) == 197
    _check(ext, "cpu", 2e-5)


def test_checkpoint_keys_are_dinos(tmp_path):
    """A DINO checkpoint (its key names and shapes) loads strictly; a `head.*` / `module.` prefixed one too."""
    src = VitExtractor("dino_vits16", device="cpu", seed=7)
    sd = src.model.state_dict()
    for k in ("cls_token", "pos_embed", "patch_embed.proj.weight", "blocks.0.norm1.weight", "blocks.11.attn.qkv.bias",
              "blocks.3.attn.proj.weight", "blocks.5.mlp.fc1.weight", "blocks.5.mlp.fc2.bias", "norm.weight"):
        assert k in sd, k
    assert sd["patch_embed.proj.weight"].shape == (384, 3, 16, 16) and sd["pos_embed"].shape == (1, 197, 384)
    path = tmp_path / "dino_deitsmall16_pretrain.pth"
    torch.save({"module." + k: v for k, v in sd.items()}, path)
    dst = VitExtractor("dino_vits16", device="cpu", weights=str(path))
    assert dst.pretrained
    x = _inputs()[0]
    assert torch.equal(dst.get_vit_attn_feat(x)["cls_"], src.get_vit_attn_feat(x)["cls_"])
    with pytest.raises(ValueError):
        VitExtractor("resnet50", device="cpu")


@pytest.mark.gpu
def test_extractor_matches_reference_gpu():
    gold = load_golden("dino_vits16")
    ext = VitExtractor("dino_vits16", device="cuda:0", seed=int(gold["seed"]))
    _check(ext, "cuda:0", 5e-4)
