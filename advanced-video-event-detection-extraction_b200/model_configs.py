"""Model geometries == open_clip/model_configs/{ViT-B-32,ViT-L-14}.json.
`pretrained="openai"` (src/utils/config.py:25-26) implies QuickGELU."""
from __future__ import annotations

from dataclasses import dataclass

from . import capi


@dataclass(frozen=True)
class ModelConfig:
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

    @property
    def patch_k(self) -> int:
        """K of the patch-embed GEMM: 3*P*P rounded up to a multiple of 64 (zero padded)."""
        return (3 * self.patch * self.patch + 63) // 64 * 64

    def flops_per_image(self) -> float:
        """Algorithmic FLOP of one image through the vision tower (MAC = 2), SURVEY.md section 8(d)."""
        t, d, f, g2 = self.tokens, self.width, self.mlp_dim, self.grid * self.grid
        per_layer = 2 * t * d * 3 * d + 2 * 2 * t * t * d + 2 * t * d * d + 2 * 2 * t * d * f
        return 2.0 * g2 * 3 * self.patch * self.patch * d + self.layers * per_layer + 2.0 * d * self.embed_dim


MODEL_CONFIGS = {
    "ViT-B-32": ModelConfig("ViT-B-32", 512, 224, 32, 768, 12, 12, 77, 49408, 512, 8, 12),
    "ViT-L-14": ModelConfig("ViT-L-14", 768, 224, 14, 1024, 24, 16, 77, 49408, 768, 12, 12),
    # small geometry used by unit tests (same code paths, seconds on any machine)
    "ViT-tiny-test": ModelConfig("ViT-tiny-test", 64, 64, 32, 128, 2, 2, 16, 512, 64, 1, 2),
}


def to_capi_config(c: ModelConfig) -> "capi.Config":
    return capi.Config(c.image_size, c.patch, c.width, c.layers, c.heads, c.mlp_dim, c.embed_dim,
                       0 if c.quick_gelu else 1, c.ln_eps, c.text_ctx, c.text_vocab, c.text_width, c.text_heads,
                       c.text_layers, c.text_mlp_dim)
