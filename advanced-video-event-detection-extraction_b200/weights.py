"""Weights in open_clip's native state-dict layout (key names: SURVEY.md Appendix A)."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from .model_configs import ModelConfig


def expected_shapes(cfg: ModelConfig) -> "OrderedDict[str, tuple]":
    W, TW, E = cfg.width, cfg.text_width, cfg.embed_dim
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["visual.class_embedding"] = (W,)
    s["visual.positional_embedding"] = (cfg.tokens, W)
    s["visual.conv1.weight"] = (W, 3, cfg.patch, cfg.patch)
    s["visual.ln_pre.weight"] = (W,)
    s["visual.ln_pre.bias"] = (W,)

    def block(prefix, width, layers):
        for i in range(layers):
            p = f"{prefix}.resblocks.{i}"
            s[f"{p}.ln_1.weight"] = (width,)
            s[f"{p}.ln_1.bias"] = (width,)
            s[f"{p}.attn.in_proj_weight"] = (3 * width, width)
            s[f"{p}.attn.in_proj_bias"] = (3 * width,)
            s[f"{p}.attn.out_proj.weight"] = (width, width)
            s[f"{p}.attn.out_proj.bias"] = (width,)
            s[f"{p}.ln_2.weight"] = (width,)
            s[f"{p}.ln_2.bias"] = (width,)
            s[f"{p}.mlp.c_fc.weight"] = (4 * width, width)
            s[f"{p}.mlp.c_fc.bias"] = (4 * width,)
            s[f"{p}.mlp.c_proj.weight"] = (width, 4 * width)
            s[f"{p}.mlp.c_proj.bias"] = (width,)

    block("visual.transformer", W, cfg.layers)
    s["visual.ln_post.weight"] = (W,)
    s["visual.ln_post.bias"] = (W,)
    s["visual.proj"] = (W, E)
    s["token_embedding.weight"] = (cfg.text_vocab, TW)
    s["positional_embedding"] = (cfg.text_ctx, TW)
    block("transformer", TW, cfg.text_layers)
    s["ln_final.weight"] = (TW,)
    s["ln_final.bias"] = (TW,)
    s["text_projection"] = (TW, E)
    return s


def random_state_dict(cfg: ModelConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random init with CLIP-like scales (used when no checkpoint is available offline, and by bench.py).
    Matrices ~ N(0, fan_in^-1/2), LayerNorm gains 1 +- 0.1, small biases."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in expected_shapes(cfg).items():
        if name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name.endswith("ln_pre.weight") \
                or name.endswith("ln_post.weight") or name.endswith("ln_final.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name == "visual.conv1.weight":
            t = torch.randn(shape, generator=g) * (3 * cfg.patch * cfg.patch) ** -0.5
        elif name in ("visual.proj", "text_projection"):
            t = torch.randn(shape, generator=g) * shape[0] ** -0.5
        elif len(shape) == 2 and name.endswith("weight") and "embedding" not in name:
            depth = cfg.layers if name.startswith("visual") else cfg.text_layers
            scale = shape[1] ** -0.5
            if "out_proj" in name or "c_proj" in name:
                scale *= (2 * depth) ** -0.5
            t = torch.randn(shape, generator=g) * scale
        elif name == "token_embedding.weight":
            t = torch.randn(shape, generator=g) * 0.08
        elif name == "positional_embedding":
            t = torch.randn(shape, generator=g) * 0.04
        else:  # class / positional embeddings of the vision tower
            t = torch.randn(shape, generator=g) * cfg.width ** -0.5
        sd[name] = t.to(torch.float32)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return sd


def load_checkpoint(path: str) -> "OrderedDict[str, torch.Tensor]":
    """A checkpoint saved in open_clip's native CLIP layout (torch.save(model.state_dict()))."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if "state_dict" in sd:
        sd = sd["state_dict"]
    return OrderedDict((k[7:] if k.startswith("module.") else k, v.float()) for k, v in sd.items())
