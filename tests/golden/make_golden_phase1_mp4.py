"""End-to-end golden on a REAL video file: the reference's own, unmodified Phase1MVP.process_video
(/root/reference/src/pipeline/phase1_mvp.py:36-163) -- OpenCV decode with seek + read per sampled frame, the <= 512 x 512
INTER_AREA shrink, sliding windows, one embedded middle frame per window on the seeded fp32 CLIP restatement, np.dot,
argsort, threshold -- on an mp4 written by tests/synth.py::write_frames_video (96 structured 640x360 frames, 8 fps).
Stored: the result list at threshold -1, the per-window similarities and timestamps from debug mode, and the result
list at a threshold that keeps three hits.

  python tests/golden/make_golden_phase1_mp4.py   (needs /root/reference; tests/golden/phase1_mp4.json is committed)
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import make_golden as mg  # noqa: E402
from synth import structured_frames, write_frames_video  # noqa: E402

N_FRAMES, H, W, FPS, SEED, QUERY, TOP_K = 96, 360, 640, 8.0, 31, "red car driving", 5


def main():
    import torch

    torch.set_num_threads(os.cpu_count() or 8)
    config = mg.import_reference("ViT-B-32")
    import src.services  # noqa: F401
    from src.pipeline.phase1_mvp import Phase1MVP

    path = write_frames_video(os.path.join(tempfile.mkdtemp(prefix="b200clip_p1_"), "clip.mp4"),
                              structured_frames(N_FRAMES, H, W, seed=SEED), FPS)
    config.settings.CONFIDENCE_THRESHOLD = -1.0
    p1 = Phase1MVP()
    results, debug = p1.process_video(path, QUERY, top_k=TOP_K, debug_mode=True)
    sims = [float(d["similarity"]) for d in debug]
    thr = sorted(sims, reverse=True)[3] + 1e-4 if len(sims) > 3 else -1.0        # keeps exactly three hits
    config.settings.CONFIDENCE_THRESHOLD = thr
    kept = Phase1MVP().process_video(path, QUERY, top_k=TOP_K, debug_mode=False)
    out = {"n_frames": N_FRAMES, "h": H, "w": W, "fps": FPS, "seed": SEED, "query": QUERY, "top_k": TOP_K,
           "results": results, "similarities": sims, "window_timestamps": [float(d["timestamp"]) for d in debug],
           "frame_shapes": [list(d["frame_shape"]) for d in debug], "threshold": thr, "results_thresholded": kept}
    with open(os.path.join(HERE, "phase1_mp4.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote phase1_mp4.json:", len(sims), "windows; top-5", [r["window_index"] for r in results], "kept", len(kept))


if __name__ == "__main__":
    main()
