"""Byte-level BPE tokenizer with CLIP's framing (open_clip `SimpleTokenizer`, call site
/root/reference/src/models/openclip_model.py:82,205).  Needs CLIP's merges file `bpe_simple_vocab_16e6.txt.gz`
(not available offline in this environment: point B200CLIP_BPE_VOCAB at it).  Algorithm as published with CLIP:
bytes -> printable unicode table, 256 + 256 '</w>' base tokens, merges[1 : 49152-256-2+1], '<|startoftext|>' /
'<|endoftext|>' appended; text is whitespace-collapsed and lower-cased, split with CLIP's regex, each word BPE-merged
greedily by merge rank; output = [SOT] + ids[:ctx-2] + [EOT], zero padded.  Of the original's ftfy/html clean-up, the
html unescape and ftfy's default NFC normalisation are kept (ftfy itself is not installed: mojibake repair is
omitted; well-formed UTF-8 text is unaffected).  The algorithm is cross-checked against an independent implementation
(transformers.CLIPTokenizer on the Rust `tokenizers` BPE) over a synthetic merges file in
tests/test_bpe_tokenizer.py; the REAL 49 152-entry vocabulary is not available offline, so ids for real CLIP
checkpoints are untested."""
from __future__ import annotations

import gzip
import html
import unicodedata
from functools import lru_cache

import regex as re   # \p{L} / \p{N} classes, as in CLIP's simple_tokenizer.py

import torch


@lru_cache()
def bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + \
        list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(2 ** 8):
        if b not in bs:
            bs.append(b)
            cs.append(2 ** 8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def get_pairs(word):
    pairs = set()
    prev = word[0]
    for ch in word[1:]:
        pairs.add((prev, ch))
        prev = ch
    return pairs


class SimpleTokenizer:
    def __init__(self, bpe_path: str, context_length: int = 77):
        self.context_length = context_length
        self.byte_encoder = bytes_to_unicode()
        opener = gzip.open if bpe_path.endswith(".gz") else open
        with opener(bpe_path, "rt", encoding="utf-8") as f:
            merges = f.read().split("\n")
        merges = merges[1:49152 - 256 - 2 + 1]
        merges = [tuple(m.split()) for m in merges if m.strip()]
        vocab = list(self.byte_encoder.values())
        vocab = vocab + [v + "</w>" for v in vocab]
        for m in merges:
            vocab.append("".join(m))
        vocab.extend(["<|startoftext|>", "<|endoftext|>"])
        self.encoder = dict(zip(vocab, range(len(vocab))))
        self.bpe_ranks = dict(zip(merges, range(len(merges))))
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pat = re.compile(
            r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", re.IGNORECASE)
        self.sot = self.encoder["<|startoftext|>"]
        self.eot = self.encoder["<|endoftext|>"]

    def bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = get_pairs(word) if len(word) > 1 else set()
        if not pairs:
            return token + "</w>"
        while True:
            bigram = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if bigram not in self.bpe_ranks:
                break
            first, second = bigram
            new_word, i = [], 0
            while i < len(word):
                try:
                    j = word.index(first, i)
                    new_word.extend(word[i:j])
                    i = j
                except ValueError:
                    new_word.extend(word[i:])
                    break
                if word[i] == first and i < len(word) - 1 and word[i + 1] == second:
                    new_word.append(first + second)
                    i += 2
                else:
                    new_word.append(word[i])
                    i += 1
            word = tuple(new_word)
            if len(word) == 1:
                break
            pairs = get_pairs(word)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text: str):
        text = unicodedata.normalize("NFC", html.unescape(html.unescape(text)))
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids = []
        for token in re.findall(self.pat, text):
            token = "".join(self.byte_encoder[b] for b in token.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(token).split(" "))
        return ids

    def __call__(self, texts) -> torch.Tensor:
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), self.context_length, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [self.sot] + self.encode(t)[: self.context_length - 2] + [self.eot]
            out[i, : len(ids)] = torch.tensor(ids)
        return out
