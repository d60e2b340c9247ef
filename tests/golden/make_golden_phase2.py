"""Golden vectors for the Phase-2 hand-off ("phase 1 is asked for 2*top_k candidates and BLIP
re-ranks 0.7*clip + 0.3*caption"), produced by the REFERENCE'S OWN Phase2Reranker.process_video
(/root/reference/src/pipeline/phase2_reranker.py:31-90) in this container, with its three collaborators replaced by
deterministic stubs so that the fixture pins the hand-off itself: how many candidates phase 1 is asked for, which frame
of which window is captioned, the score combination, the result keys, the stable descending sort and the truncation.

  * phase 1: returns a prepared hit list truncated to the requested count (and a debug tuple in debug mode);
  * frame extractor: the reference's own FrameExtractor window arithmetic on tiny synthetic frames whose first pixel
    encodes the frame index;
  * BLIP: caption = "frame <index>", caption/query similarity from a seeded table.

Run:  python tests/golden/make_golden_phase2.py   (needs /root/reference; tests/golden/phase2_handoff.json is committed)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402


def tiny_frames(n):
    f = np.zeros((n, 2, 2, 3), np.uint8)
    f[:, 0, 0, 0] = np.arange(n) % 256
    f[:, 0, 0, 1] = np.arange(n) // 256
    return f


def frame_index(frame) -> int:
    return int(frame[0, 0, 0]) + 256 * int(frame[0, 0, 1])


def main():
    mg.import_reference("ViT-B-32")
    import src.services  # noqa: F401
    from src.pipeline.phase2_reranker import Phase2Reranker
    from src.services.frame_extractor import FrameExtractor

    rng = np.random.default_rng(77)
    cases = []
    for case in range(16):
        n_frames = int(rng.choice([40, 80, 123, 10]))
        fx = FrameExtractor()
        frames = tiny_frames(n_frames)
        stamps = [round(i / 4.0, 4) for i in range(n_frames)]
        windows, window_ts = fx.create_sliding_windows(frames, stamps)
        m = len(window_ts)
        conf = np.round(rng.uniform(0.2, 0.6, m), 3)
        order = np.argsort(conf)[::-1]
        hits = [{"timestamp": window_ts[int(i)], "confidence": float(conf[int(i)]), "phase": "phase1_mvp", "window_index": int(i)}
                for i in order]
        cap_table = np.round(rng.uniform(0.0, 1.0, n_frames), 3)
        if case % 3 == 0 and m > 3:               # ties in the combined score: equal clip scores, equal caption scores
            hits[1]["confidence"] = hits[0]["confidence"]
            cap_table[:] = 0.5
        top_k = [None, 3, 5, 1][case % 4]
        debug = case % 5 == 4
        if case == 7:
            hits = []                              # phase 1 found nothing
        asked = []

        class P1:
            def process_video(self, video_path, query, top_k=None, debug_mode=None):
                asked.append(top_k)
                res = [dict(h) for h in hits[:top_k]]
                return (res, [{"window_index": i} for i in range(m)]) if debug_mode else res

        class FX:
            def extract_frames(self, video_path):
                return frames, stamps

            def create_sliding_windows(self, fr, ts):
                return fx.create_sliding_windows(fr, ts)

        class Blip:
            def generate_caption(self, frame):
                return f"frame {frame_index(frame)}"

            def compute_text_similarity(self, caption, query):
                return float(cap_table[int(caption.split()[1])])

        r = object.__new__(Phase2Reranker)
        r.phase1, r.frame_extractor, r.blip_model, r.lazy_load = P1(), FX(), Blip(), True
        out = r.process_video("video.mp4", "red car", top_k, debug_mode=debug)
        cases.append({"n_frames": n_frames, "fps": 4.0, "hits": hits, "caption_table": cap_table.tolist(), "top_k": top_k,
                      "debug": debug, "asked_phase1_for": asked[0], "output": out})
    with open(os.path.join(HERE, "phase2_handoff.json"), "w") as f:
        json.dump({"cases": cases}, f, indent=0)
    print("wrote phase2_handoff.json:", len(cases), "cases")


if __name__ == "__main__":
    main()
