"""ORACLE (test infrastructure, not product): CPU fp32 PyTorch restatement of open_clip's CLIP.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path never does (it fails loudly if libb200clip.so is missing).

What it restates
----------------
The reference's hot-path arithmetic lives in the third-party package `open_clip_torch>=2.20.0`
(/root/reference/requirements.txt:5, not vendored, not installed here, no lock file).  The reference
calls it at src/models/openclip_model.py:77-83 (create_model_and_transforms / get_tokenizer / eval),
:177,196 (model.encode_image) and :205-208 (tokenizer, model.encode_text).  This file restates the
published algorithm of open_clip's `CLIP` / `VisionTransformer` / `TextTransformer` for the
ViT-B-32 and ViT-L-14 configs with `pretrained="openai"` semantics (QuickGELU, fp32), using open_clip's
state-dict key names:

  vision : conv1(P x P, stride P, no bias) -> [cls ; patches] + positional_embedding -> ln_pre
           -> L x { x += MHA(ln_1(x)) ; x += c_proj(act(c_fc(ln_2(x)))) } -> ln_post(x[:,0]) @ proj
  text   : token_embedding[ids] + positional_embedding -> same blocks with an additive causal mask
           -> ln_final -> row at argmax(ids) (EOT has the largest id) @ text_projection

Pinning: the reference's tests hold no golden vector for this path, so the pins are
(1) tests/golden/*.npz produced by running the reference's own unmodified wrapper
(src/models/openclip_model.py::OpenCLIPModel) on top of this restatement (tests/golden/make_golden.py), and
(2) an independent cross-check against `transformers.CLIPModel` with copied weights
(tests/test_oracle_clip.py).  String -> BPE tokenisation fidelity is unverified (no vocab file offline).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class CLIPConfig:
    name: str
    embed_dim: int
    image_size: int
    patch: int
    width: int
    layers: int
    heads: int
    text_ctx: int = 77
    text_vocab: int = 49408
    text_width: int = 512
    text_heads: int = 8
    text_layers: int = 12
    quick_gelu: bool = True
    ln_eps: float = 1e-5

    @property
    def mlp_dim(self) -> int:
        return 4 * self.width

    @property
    def text_mlp_dim(self) -> int:
        return 4 * self.text_width

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1


# open_clip/model_configs/ViT-B-32.json and ViT-L-14.json
CONFIGS = {
    "ViT-B-32": CLIPConfig("ViT-B-32", 512, 224, 32, 768, 12, 12, 77, 49408, 512, 8, 12),
    "ViT-L-14": CLIPConfig("ViT-L-14", 768, 224, 14, 1024, 24, 16, 77, 49408, 768, 12, 12),
    # a tiny geometry for fast CPU tests of the same code paths (not an open_clip config)
    "ViT-tiny-test": CLIPConfig("ViT-tiny-test", 64, 64, 32, 128, 2, 2, 16, 512, 64, 1, 2),
}


def init_state_dict(cfg: CLIPConfig, seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random init in open_clip's state-dict layout.

    Follows open_clip's `CLIP.init_parameters` scales (attn std = width^-0.5, proj std =
    width^-0.5 * (2 layers)^-0.5, fc std = (2 width)^-0.5; embeddings 0.02 / 0.01) with non-trivial
    LayerNorm affine and biases so that every term of the forward is exercised.  `gain` widens the
    projection matrices so that image x text scores spread beyond the parity tolerance.
    Deterministic across machines for a fixed torch version (CPU generator).
    """
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    W, L = cfg.width, cfg.layers
    scale = W ** -0.5
    sd["visual.class_embedding"] = rn(W, std=scale)
    sd["visual.positional_embedding"] = rn(cfg.tokens, W, std=scale)
    sd["visual.conv1.weight"] = rn(W, 3, cfg.patch, cfg.patch, std=(3 * cfg.patch * cfg.patch) ** -0.5)
    sd["visual.ln_pre.weight"] = 1.0 + rn(W, std=0.1)
    sd["visual.ln_pre.bias"] = rn(W, std=0.1)

    def block(prefix: str, width: int, layers: int):
        attn_std = width ** -0.5
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        fc_std = (2 * width) ** -0.5
        for i in range(layers):
            p = f"{prefix}.resblocks.{i}"
            sd[f"{p}.ln_1.weight"] = 1.0 + rn(width, std=0.1)
            sd[f"{p}.ln_1.bias"] = rn(width, std=0.1)
            sd[f"{p}.attn.in_proj_weight"] = rn(3 * width, width, std=attn_std * gain)
            sd[f"{p}.attn.in_proj_bias"] = rn(3 * width, std=0.02)
            sd[f"{p}.attn.out_proj.weight"] = rn(width, width, std=proj_std * gain)
            sd[f"{p}.attn.out_proj.bias"] = rn(width, std=0.02)
            sd[f"{p}.ln_2.weight"] = 1.0 + rn(width, std=0.1)
            sd[f"{p}.ln_2.bias"] = rn(width, std=0.1)
            sd[f"{p}.mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std * gain)
            sd[f"{p}.mlp.c_fc.bias"] = rn(4 * width, std=0.02)
            sd[f"{p}.mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std * gain)
            sd[f"{p}.mlp.c_proj.bias"] = rn(width, std=0.02)

    block("visual.transformer", W, L)
    sd["visual.ln_post.weight"] = 1.0 + rn(W, std=0.1)
    sd["visual.ln_post.bias"] = rn(W, std=0.1)
    sd["visual.proj"] = rn(W, cfg.embed_dim, std=scale)

    TW = cfg.text_width
    sd["token_embedding.weight"] = rn(cfg.text_vocab, TW, std=0.02 * gain * 4)
    sd["positional_embedding"] = rn(cfg.text_ctx, TW, std=0.01 * gain * 4)
    block("transformer", TW, cfg.text_layers)
    sd["ln_final.weight"] = 1.0 + rn(TW, std=0.1)
    sd["ln_final.bias"] = rn(TW, std=0.1)
    sd["text_projection"] = rn(TW, cfg.embed_dim, std=TW ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return sd


def _act(x: torch.Tensor, quick: bool) -> torch.Tensor:
    # open_clip QuickGELU: x * sigmoid(1.702 x); otherwise nn.GELU() (erf)
    return x * torch.sigmoid(1.702 * x) if quick else F.gelu(x)


def _mha(x: torch.Tensor, sd, p: str, heads: int, mask: torch.Tensor | None) -> torch.Tensor:
    """nn.MultiheadAttention(width, heads) forward on [B, T, W] with packed q,k,v in_proj."""
    B, T, W = x.shape
    hd = W // heads
    qkv = F.linear(x, sd[f"{p}.attn.in_proj_weight"], sd[f"{p}.attn.in_proj_bias"])
    q, k, v = qkv.split(W, dim=-1)
    q = q.view(B, T, heads, hd).transpose(1, 2)
    k = k.view(B, T, heads, hd).transpose(1, 2)
    v = v.view(B, T, heads, hd).transpose(1, 2)
    s = (q * (hd ** -0.5)) @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    a = torch.softmax(s, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, T, W)
    return F.linear(o, sd[f"{p}.attn.out_proj.weight"], sd[f"{p}.attn.out_proj.bias"])


def _blocks(x: torch.Tensor, sd, prefix: str, layers: int, heads: int, quick: bool, eps: float,
            mask: torch.Tensor | None) -> torch.Tensor:
    W = x.shape[-1]
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}"
        h = F.layer_norm(x, (W,), sd[f"{p}.ln_1.weight"], sd[f"{p}.ln_1.bias"], eps)
        x = x + _mha(h, sd, p, heads, mask)
        h = F.layer_norm(x, (W,), sd[f"{p}.ln_2.weight"], sd[f"{p}.ln_2.bias"], eps)
        h = _act(F.linear(h, sd[f"{p}.mlp.c_fc.weight"], sd[f"{p}.mlp.c_fc.bias"]), quick)
        x = x + F.linear(h, sd[f"{p}.mlp.c_proj.weight"], sd[f"{p}.mlp.c_proj.bias"])
    return x


class CLIPRef:
    """Object with the slice of the open_clip model surface the reference wrapper touches:
    .eval(), .encode_image(FloatTensor[B,3,S,S]) -> [B,E], .encode_text(LongTensor[Q,ctx]) -> [Q,E]
    (both un-normalised), .state_dict()."""

    def __init__(self, cfg: CLIPConfig, sd):
        self.cfg = cfg
        self.sd = {k: v.detach().to(torch.float32) for k, v in sd.items()}

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def state_dict(self):
        return self.sd

    @torch.no_grad()
    def encode_image(self, image: torch.Tensor, upto: str | None = None) -> torch.Tensor:
        cfg, sd = self.cfg, self.sd
        x = F.conv2d(image.to(torch.float32), sd["visual.conv1.weight"], stride=cfg.patch)  # [B, W, g, g]
        B = x.shape[0]
        x = x.reshape(B, cfg.width, -1).permute(0, 2, 1)                                     # [B, g*g, W]
        cls = sd["visual.class_embedding"].expand(B, 1, cfg.width)
        x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"]
        if upto == "embed":
            return x
        x = F.layer_norm(x, (cfg.width,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"], cfg.ln_eps)
        if upto == "ln_pre":
            return x
        x = _blocks(x, sd, "visual.transformer", cfg.layers, cfg.heads, cfg.quick_gelu, cfg.ln_eps, None)
        if upto == "blocks":
            return x
        pooled = F.layer_norm(x[:, 0], (cfg.width,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"],
                              cfg.ln_eps)
        return pooled @ sd["visual.proj"]

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        cfg, sd = self.cfg, self.sd
        T = text.shape[1]
        x = sd["token_embedding.weight"][text] + sd["positional_embedding"][:T]
        mask = torch.full((T, T), float("-inf")).triu_(1)
        x = _blocks(x, sd, "transformer", cfg.text_layers, cfg.text_heads, cfg.quick_gelu, cfg.ln_eps, mask)
        x = F.layer_norm(x, (cfg.text_width,), sd["ln_final.weight"], sd["ln_final.bias"], cfg.ln_eps)
        pooled = x[torch.arange(x.shape[0]), text.argmax(dim=-1)]
        return pooled @ sd["text_projection"]


def synthetic_tokenize(texts, ctx: int = 77, vocab: int = 49408) -> torch.Tensor:
    """Deterministic stand-in for open_clip's SimpleTokenizer (its BPE vocab file is not available
    offline): [SOT] + one id per whitespace-separated lower-cased word + [EOT], zero padded to ctx.
    SOT = vocab-2 (49406), EOT = vocab-1 (49407) as in CLIP, word ids are a stable hash in [1, vocab-3]
    so that argmax(ids) lands on EOT exactly as with the real tokenizer."""
    import zlib

    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), ctx, dtype=torch.long)
    sot, eot = vocab - 2, vocab - 1
    for i, t in enumerate(texts):
        words = t.lower().split()
        ids = [sot] + [1 + zlib.crc32(w.encode("utf-8")) % (vocab - 3) for w in words][: ctx - 2] + [eot]
        out[i, : len(ids)] = torch.tensor(ids)
    return out
