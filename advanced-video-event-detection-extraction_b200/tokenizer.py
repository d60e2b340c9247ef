"""Tokenizer front end.  open_clip's SimpleTokenizer needs `bpe_simple_vocab_16e6.txt.gz`, which is not
available offline: point B200CLIP_BPE_VOCAB at it and the byte-level BPE of bpe.py is used.
Without it get_tokenizer() RAISES -- a trained text tower fed with stand-in ids returns meaningless hits -- unless
B200CLIP_ALLOW_SYNTHETIC=1 explicitly opts into the deterministic stand-in below (same framing: [SOT] ids... [EOT],
zero padded to 77, EOT the largest id so that argmax pooling lands on it).  Tests, bench.py and smoke() opt in: they
run seeded random weights, for which any fixed string -> id map is as good as another."""
from __future__ import annotations

import os
import zlib

import torch

from .model_configs import MODEL_CONFIGS


class HashTokenizer:
    def __init__(self, ctx: int = 77, vocab: int = 49408):
        self.ctx, self.vocab = ctx, vocab
        self.sot, self.eot = vocab - 2, vocab - 1

    def __call__(self, texts) -> torch.Tensor:
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), self.ctx, dtype=torch.long)
        for i, t in enumerate(texts):
            words = t.lower().split()
            ids = [self.sot] + [1 + zlib.crc32(w.encode("utf-8")) % (self.vocab - 3) for w in words][: self.ctx - 2] \
                + [self.eot]
            out[i, : len(ids)] = torch.tensor(ids)
        return out


def get_tokenizer(model_name: str):
    """open_clip.get_tokenizer(name) -> callable(list[str]) -> LongTensor[Q, ctx]
    (reference call site: src/models/openclip_model.py:82)."""
    cfg = MODEL_CONFIGS[model_name.replace("/", "-")]
    vocab_path = os.environ.get("B200CLIP_BPE_VOCAB")
    if vocab_path:
        if not os.path.exists(vocab_path):
            raise FileNotFoundError(f"B200CLIP_BPE_VOCAB={vocab_path} does not exist")
        from .bpe import SimpleTokenizer

        return SimpleTokenizer(vocab_path, cfg.text_ctx)
    if not allow_synthetic():
        raise RuntimeError(
            "no BPE vocabulary: set B200CLIP_BPE_VOCAB to CLIP's bpe_simple_vocab_16e6.txt.gz, or set "
            "B200CLIP_ALLOW_SYNTHETIC=1 to accept the stand-in HashTokenizer (only meaningful with synthetic weights)")
    return HashTokenizer(cfg.text_ctx, cfg.text_vocab)


def allow_synthetic() -> bool:
    """Explicit opt-in to stand-ins (hash tokenizer, seeded random weights) -- never a silent default."""
    return os.environ.get("B200CLIP_ALLOW_SYNTHETIC", "0") not in ("", "0")
