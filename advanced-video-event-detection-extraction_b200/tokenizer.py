"""Tokenizer front end.  open_clip's SimpleTokenizer needs `bpe_simple_vocab_16e6.txt.gz`, which is not
available offline; if a vocab file is supplied (env B200CLIP_BPE_VOCAB) a byte-level BPE over
it is used, otherwise a deterministic stand-in with the same framing ([SOT] ids... [EOT], zero padded to 77, EOT
the largest id so that argmax pooling lands on it).  String -> BPE fidelity is therefore unverified here."""
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
    if vocab_path and os.path.exists(vocab_path):
        from .bpe import SimpleTokenizer

        return SimpleTokenizer(vocab_path, cfg.text_ctx)
    return HashTokenizer(cfg.text_ctx, cfg.text_vocab)
