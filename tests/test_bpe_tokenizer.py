"""String -> token-id parity of b200clip.bpe.SimpleTokenizer (open_clip's tokenizer; reference call sites
/root/reference/src/models/openclip_model.py:82,205) against an INDEPENDENT implementation of the same algorithm:
transformers.CLIPTokenizer (Rust `tokenizers` BPE), both built from the synthetic merges fixture written by
tests/golden/make_bpe_fixture.py.  The real 49 152-entry vocabulary is not available offline: what is pinned here is
the algorithm (normalisation, CLIP's split regex, '</w>' handling, greedy merge order, framing), not the real ids."""
import os

import pytest
import torch

from b200clip.bpe import SimpleTokenizer, bytes_to_unicode

MERGES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bpe_merges.txt")

TEXTS = [
    "a person walking across the street",
    "A Person   WALKING\tacross\nthe   street  ",                  # case + whitespace collapse
    "someone's opening the door, isn't it? they're here; we've won",
    "platform 12 at 10:45, gate 7b",                                # digits split one by one
    "shoot-out!!! (penalty) -- goal... #1",
    "l'été à münchen über straße café",                             # non-ASCII letters are \p{L}
    "naïve façade — “quoted” text…",
    "日本語 の テキスト 123",
    "x",
    "",
    "&amp; &lt;b&gt; html &amp;amp; entities",                      # open_clip unescapes html twice
    "thethethe walkingwalking personperson",
    "zzzzqqqq unknownwordswithnomerges",
    "a " * 120,                                                     # longer than the 77-token context
    "the quick brown fox jumps over the lazy dog " * 6,
]


@pytest.fixture(scope="module")
def ours():
    return SimpleTokenizer(MERGES, 77)


@pytest.fixture(scope="module")
def hf():
    from transformers import CLIPTokenizer

    with open(MERGES, encoding="utf-8") as f:
        merges = [tuple(ln.split()) for ln in f.read().split("\n")[1:] if ln.strip()]
    base = list(bytes_to_unicode().values())
    vocab = base + [v + "</w>" for v in base] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
    return CLIPTokenizer(vocab={t: i for i, t in enumerate(vocab)}, merges=merges)


def hf_clean(text: str) -> str:
    import html

    return html.unescape(html.unescape(text))        # the html clean-up is open_clip's, not HF's


def test_vocabulary_layout(ours):
    n_merges = sum(1 for ln in open(MERGES, encoding="utf-8").read().split("\n")[1:] if ln.strip())
    assert len(ours.encoder) == 512 + n_merges + 2
    assert ours.sot == len(ours.encoder) - 2 and ours.eot == len(ours.encoder) - 1     # EOT is the largest id
    assert ours.encoder["a</w>"] == 256 + list(bytes_to_unicode().values()).index("a")


@pytest.mark.parametrize("text", TEXTS)
def test_ids_match_the_independent_bpe(ours, hf, text):
    want = hf(hf_clean(text), truncation=True, max_length=77)["input_ids"]
    row = ours([text])[0]
    n = len(want)
    assert row[:n].tolist() == want
    assert int(row[n:].abs().sum()) == 0                            # zero padded (open_clip), HF pads with EOT


def test_framing_truncation_and_argmax_pooling(ours):
    rows = ours(["a " * 120, "hello there", ""])
    assert rows.shape == (3, 77) and rows.dtype == torch.long
    assert rows[0, 0] == ours.sot and rows[0, 76] == ours.eot       # truncated to ctx with EOT forced last
    assert rows[1, 0] == ours.sot
    # EOT = max id, so open_clip's text.argmax(dim=-1) pooling lands on it (first occurrence)
    for r in rows:
        assert r[int(r.argmax())] == ours.eot
    assert rows[2, :2].tolist() == [ours.sot, ours.eot]


def test_str_input_and_cache(ours):
    a = ours("a person walking")
    b = ours(["a person walking"])
    assert torch.equal(a, b)


def test_get_tokenizer_fails_loudly_without_vocab(monkeypatch):
    from b200clip import tokenizer

    monkeypatch.delenv("B200CLIP_BPE_VOCAB", raising=False)
    monkeypatch.setenv("B200CLIP_ALLOW_SYNTHETIC", "0")
    with pytest.raises(RuntimeError, match="B200CLIP_BPE_VOCAB"):
        tokenizer.get_tokenizer("ViT-B-32")
    monkeypatch.setenv("B200CLIP_BPE_VOCAB", "/nonexistent/vocab.txt.gz")
    with pytest.raises(FileNotFoundError):
        tokenizer.get_tokenizer("ViT-B-32")
    monkeypatch.setenv("B200CLIP_BPE_VOCAB", MERGES)
    tok = tokenizer.get_tokenizer("ViT-B-32")
    assert isinstance(tok, SimpleTokenizer) and tok(["a dog"]).shape == (1, 77)
    monkeypatch.delenv("B200CLIP_BPE_VOCAB")
    monkeypatch.setenv("B200CLIP_ALLOW_SYNTHETIC", "1")
    assert isinstance(tokenizer.get_tokenizer("ViT-B-32"), tokenizer.HashTokenizer)
