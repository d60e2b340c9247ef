#!/usr/bin/env python
"""Writes tests/golden/bpe_merges.txt: a small byte-level BPE merges file in the format of CLIP's
`bpe_simple_vocab_16e6.txt` (header line, then one "left right" merge per line), LEARNED here from the fixed corpus
below with the textbook BPE procedure (most frequent adjacent pair first, ties -> lexicographically smallest).

The real 49 152-entry vocabulary cannot be downloaded in this environment; this fixture lets tests/test_bpe_tokenizer.py
run b200clip's SimpleTokenizer (open_clip's tokenizer, reference call sites src/models/openclip_model.py:82,205) against
an INDEPENDENT implementation of the same algorithm -- transformers.CLIPTokenizer on the Rust `tokenizers` BPE -- built
from the same merges.  Deterministic: re-running reproduces the committed file byte for byte.

    python tests/golden/make_bpe_fixture.py
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from b200clip.bpe import bytes_to_unicode  # noqa: E402

CORPUS = """
a person walking across the street at night . a man in a red shirt running through the park
two people shaking hands near a white car , a woman carrying a blue backpack enters the building
someone opening the door of a black truck ; a dog chasing a ball on the grass
the camera shows a crowded train station with travellers waiting on platform 12
a cyclist wearing a yellow helmet stops at the traffic light and waits for the signal
children playing football in the school yard while it 's raining heavily
a delivery driver drops a package at the front door and walks back to the van
person falling down the stairs ! security guard checking bags at the entrance
a white van parked in front of the shop for 45 minutes , nobody 's inside
people dancing at a wedding party , the bride 's dress is white and the groom wears a dark suit
find the moment when the goalkeeper catches the ball during the penalty shoot-out
show me when the speaker points at the whiteboard and explains the diagram
un homme traverse la rue près du café , l'été à münchen ; über straße
""" * 3


def main():
    enc = bytes_to_unicode()
    words = collections.Counter()
    for w in CORPUS.lower().split():
        sym = tuple(enc[b] for b in w.encode("utf-8"))
        words[sym[:-1] + (sym[-1] + "</w>",)] += 1
    merges = []
    for _ in range(420):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        top = max(pairs.values())
        best = min(p for p, c in pairs.items() if c == top)
        merges.append(best)
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1])
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        words = new
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpe_merges.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("#version: 0.2 (synthetic fixture, see make_bpe_fixture.py)\n")
        for a, b in merges:
            f.write(f"{a} {b}\n")
    print(f"wrote {len(merges)} merges to {path}")


if __name__ == "__main__":
    main()
