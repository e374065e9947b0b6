"""DINO ViT feature provider -- drop-in for the reference's `models/extractor.py`.

The reference wraps `torch.hub.load('facebookresearch/dino:main', 'dino_vits16')` in forward hooks and reads, per call,
the last block's output tokens and attention probabilities (extractor.py:106-113, 204-224).  There is no network
here and the hub entry point is not needed: the ViT-S/16 (or /8, ViT-B) is small enough to state directly.  This
module keeps

  * the constructor `VitExtractor(model_name, device)` (extractor.py:27) plus an optional `weights=` path to a DINO
    checkpoint (`dino_deitsmall16_pretrain.pth` and friends: the parameter names below are the checkpoint's, so
    `load_state_dict` is strict); without weights the net is seeded random-init (trunc-normal 0.02, like the hub
    model before loading) and `.pretrained` is False;
  * `get_vit_attn_feat(x)` -> {'attn' [B,1,N], 'cls_' [B,C], 'feat' [B,N,C]} at 224 x 224 (nearest resize,
    extractor.py:204-213), `get_vit_attn_feat_noresize` (:215-224, bicubic position-embedding interpolation as in
    DINO's `interpolate_pos_encoding`), `get_vit_feature` (:183-190), `get_vit_feature_attn` (:193-201);
  * the values: block outputs are taken BEFORE the final LayerNorm (the reference hooks `model.blocks[i]`), the
    attention is the softmax of the last block averaged over heads, CLS row, patch columns.

What changes: no hooks and no per-call hook registration, one pass that materialises only what is returned (the
reference stores q/k/v, attention maps and block outputs of all 12 layers on every call), fused attention
(`scaled_dot_product_attention`) in the 11 blocks whose probabilities are not needed, everything under `no_grad`
(the reference builds an autograd graph it never uses).  PyTorch/cuBLAS is the right tool for this
row: it is a library transformer forward outside the render hot path.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_ARCH = {  # model_name -> (embed_dim, depth, heads, patch)
    "dino_vits16": (384, 12, 6, 16), "dino_vits8": (384, 12, 6, 8),
    "dino_vitb16": (768, 12, 12, 16), "dino_vitb8": (768, 12, 12, 8),
}
_MEAN, _STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim)

    def forward(self, x, want_attn=False):
        B, N, C = x.shape
        h = self.attn.heads
        qkv = self.attn.qkv(self.norm1(x)).view(B, N, 3, h, C // h).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        probs = None
        if want_attn:
            probs = ((q @ k.transpose(-2, -1)) * (C // h) ** -0.5).softmax(dim=-1)
            y = probs @ v
        else:
            y = F.scaled_dot_product_attention(q, k, v)
        x = x + self.attn.proj(y.transpose(1, 2).reshape(B, N, C))
        x = x + self.mlp.fc2(F.gelu(self.mlp.fc1(self.norm2(x))))
        return x, probs


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch):
        super().__init__()
        self.patch_size = patch
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)   # checkpoint layout [dim, 3, p, p]

    def forward(self, x):
        # non-overlapping patches -> one GEMM (no cuDNN/TF32 path: fp32 like the rest of the provider)
        B, C, H, W = x.shape
        p = self.patch_size
        x = x[:, :, :H // p * p, :W // p * p].reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5)
        return F.linear(x.reshape(B, (H // p) * (W // p), C * p * p), self.proj.weight.flatten(1), self.proj.bias)


class _ViT(nn.Module):
    def __init__(self, dim, depth, heads, patch):
        super().__init__()
        self.patch_embed = _PatchEmbed(dim, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, (224 // patch) ** 2 + 1, dim))
        self.blocks = nn.ModuleList([_Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)                            # in the checkpoint; not applied to block outputs
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def pos_encoding(self, n_h, n_w):
        """Position embedding for an n_h x n_w patch grid (bicubic, DINO's +0.1 scale-factor convention)."""
        n0 = int(math.sqrt(self.pos_embed.shape[1] - 1))
        if n_h == n0 and n_w == n0:
            return self.pos_embed
        grid = self.pos_embed[:, 1:].reshape(1, n0, n0, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, scale_factor=((n_h + 0.1) / n0, (n_w + 0.1) / n0), mode="bicubic")
        assert grid.shape[-2:] == (n_h, n_w)
        return torch.cat([self.pos_embed[:, :1], grid.permute(0, 2, 3, 1).reshape(1, n_h * n_w, -1)], 1)

    def tokens(self, x):
        p = self.patch_embed.patch_size
        t = self.patch_embed(x)
        t = torch.cat([self.cls_token.expand(t.shape[0], -1, -1), t], 1)
        return t + self.pos_encoding(x.shape[2] // p, x.shape[3] // p)

    def last_block(self, x, want_attn):
        """Output tokens of the last block (pre-norm) and, optionally, its attention probabilities [B,h,N,N]."""
        t = self.tokens(x)
        for blk in self.blocks[:-1]:
            t, _ = blk(t)
        return self.blocks[-1](t, want_attn)


class VitExtractor(nn.Module):
    def __init__(self, model_name="dino_vits16", device="cuda", weights=None, seed=0):
        super().__init__()
        if model_name not in _ARCH:
            raise ValueError(f"unknown model {model_name!r}; known: {sorted(_ARCH)}")
        self.model_name = model_name
        dim, depth, heads, patch = _ARCH[model_name]
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(seed)
            self.model = _ViT(dim, depth, heads, patch)
        self.pretrained = weights is not None
        if weights is not None:
            sd = torch.load(weights, map_location="cpu") if isinstance(weights, (str, bytes)) or hasattr(weights, "__fspath__") else weights
            sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
            self.model.load_state_dict({k: v for k, v in sd.items() if not k.startswith("head.")}, strict=True)
        self.model.to(device).eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self.register_buffer("_mean", torch.tensor(_MEAN).view(1, 3, 1, 1).to(device), persistent=False)
        self.register_buffer("_std", torch.tensor(_STD).view(1, 3, 1, 1).to(device), persistent=False)

    # -- the reference's small accessors (extractor.py:115-148)
    def get_patch_size(self):
        return _ARCH[self.model_name][3]

    def get_head_num(self):
        return _ARCH[self.model_name][2]

    def get_embedding_dim(self):
        return _ARCH[self.model_name][0]

    def get_patch_num(self, input_img_shape):
        _, _, h, w = input_img_shape
        p = self.get_patch_size()
        return 1 + (h // p) * (w // p)

    def _norm(self, x, resize):
        if resize:
            x = F.interpolate(x, size=(224, 224))                              # nearest, as the reference
        return (x - self._mean.to(x.dtype)) / self._std.to(x.dtype)

    @torch.no_grad()
    def _attn_feat(self, x, resize):
        t, probs = self.model.last_block(self._norm(x, resize), want_attn=True)
        return {"attn": probs.mean(1).unsqueeze(1)[:, :, 0, 1:], "cls_": t[:, 0, :], "feat": t[:, 1:, :]}

    def get_vit_attn_feat(self, x):
        """x [B,3,h,w] in [0,1] -> attn [B,1,196], cls_ [B,C], feat [B,196,C] (extractor.py:204-213)."""
        return self._attn_feat(x, True)

    def get_vit_attn_feat_noresize(self, x):
        """Same at the native resolution (extractor.py:215-224); h, w multiples of the patch size."""
        return self._attn_feat(x, False)

    @torch.no_grad()
    def get_vit_feature(self, x):
        """Patch tokens of the last block at the native resolution (extractor.py:183-190)."""
        return self.model.last_block(self._norm(x, False), want_attn=False)[0][:, 1:, :]

    @torch.no_grad()
    def get_vit_feature_attn(self, x):
        """CLS token of the last block at 224 x 224 (extractor.py:193-201)."""
        return self.model.last_block(self._norm(x, True), want_attn=False)[0][:, 0, :]
