"""Golden vectors for the SURVEY.md section 8f "next" rows, produced by the REFERENCE'S OWN code in this container
(same stub recipe as make_golden.py):

  * temporal_consistency: Phase3Advanced._apply_temporal_consistency (/root/reference/src/pipeline/phase3_advanced.py
    :37-81) on seeded random hit lists, with and without explicit start_time/end_time, with ties;
  * single_stage_matching: ImageMatcher._single_stage_matching (/root/reference/src/services/image_matcher.py:980-1018)
    with `_compute_clip_similarity` replaced by a table lookup, so that the fixture pins the sort / top-k / threshold /
    tie behaviour (Python's stable sort with reverse=True: equal confidences keep frame order) and the result keys.

Run:  python tests/golden/make_golden_next.py     (needs /root/reference; tests/golden/next_rows.json is committed)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402


def main():
    mg.import_reference("ViT-B-32")
    import src.services  # noqa: F401
    from src.pipeline.phase3_advanced import Phase3Advanced
    from src.services.image_matcher import ImageMatcher

    rng = np.random.default_rng(2024)
    out = {"temporal_consistency": [], "single_stage_matching": []}

    for case in range(40):
        n = int(rng.integers(0, 14))
        span = float(rng.choice([6.0, 20.0, 60.0]))
        hits = []
        for i in range(n):
            h = {"timestamp": round(float(rng.uniform(0, span)), 2), "confidence": round(float(rng.uniform(0.2, 0.9)), 3),
                 "phase": "phase1_mvp", "window_index": i}
            if case % 3 == 1:                      # UniVTG-style explicit boundaries
                d = float(rng.uniform(0.5, 8.0))
                h["start_time"] = round(h["timestamp"] - d * float(rng.uniform(0.2, 0.8)), 2)
                h["end_time"] = round(h["start_time"] + d, 2)
            if case % 5 == 2 and i % 3 == 0 and hits:      # equal confidences and equal timestamps
                h["confidence"] = hits[-1]["confidence"]
                if i % 2 == 0:
                    h["timestamp"] = hits[-1]["timestamp"]
            hits.append(h)
        kept = Phase3Advanced._apply_temporal_consistency(None, [dict(h) for h in hits])
        out["temporal_consistency"].append({"input": hits, "output": kept})

    class _Stub:
        max_frames_per_batch = 8

        def __init__(self, table):
            self.table = table

        def _compute_clip_similarity(self, reference_image, frame):
            return float(self.table[int(frame[0, 0, 0]) + 256 * int(frame[0, 0, 1])])

    for case in range(12):
        n = int(rng.integers(1, 40))
        sims = np.round(rng.uniform(0.3, 0.95, n), 3)
        if case % 2 == 0 and n > 4:
            sims[n // 2] = sims[1]
            sims[n - 1] = sims[1]                  # three-way tie
        frames = np.zeros((n, 2, 2, 3), np.uint8)
        frames[:, 0, 0, 0] = np.arange(n) % 256
        frames[:, 0, 0, 1] = np.arange(n) // 256
        ts = [round(i / 3.0, 4) for i in range(n)]
        top_k = int(rng.integers(1, 8))
        thr = float(rng.choice([0.0, 0.5, 0.7, 0.9]))
        res = ImageMatcher._single_stage_matching(_Stub(sims), np.zeros((2, 2, 3), np.uint8), frames, ts, top_k, thr)
        out["single_stage_matching"].append({"similarities": sims.tolist(), "timestamps": ts, "top_k": top_k,
                                             "threshold": thr, "output": res})

    with open(os.path.join(HERE, "next_rows.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote next_rows.json:", len(out["temporal_consistency"]), "temporal cases,",
          len(out["single_stage_matching"]), "matching cases")


if __name__ == "__main__":
    main()
