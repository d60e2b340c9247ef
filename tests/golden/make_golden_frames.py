"""Golden for the decode-side semantics (SURVEY.md section 8 rows a1/a2 + 8f-3): the REFERENCE'S OWN
FrameExtractor.extract_frames (OpenCV path, /root/reference/src/services/frame_extractor.py:106-235: sampling, the
1000-frame cap, seek + read per index, timestamps from the decoder position, resize_frame_for_memory) and
create_sliding_windows (:237-273) on small mp4 files written by tests/synth.py::write_test_video.  Stored: timestamps,
window timestamps, and a checksum of every (shrunk) frame the reference kept.

  python tests/golden/make_golden_frames.py     (needs /root/reference; tests/golden/frame_extractor.json is committed)
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import make_golden as mg  # noqa: E402
from synth import write_test_video  # noqa: E402

CASES = [  # (frames, width, height, fps, FRAME_SAMPLE_RATE)
    (80, 96, 64, 8.0, 1), (83, 96, 64, 8.0, 3), (10, 96, 64, 8.0, 1), (16, 96, 64, 25.0, 1), (120, 640, 360, 30.0, 2),
    (1100, 32, 32, 30.0, 1),      # more than 1000 sampled frames: the cap (step = 1, first 1000)
    (2500, 32, 32, 30.0, 1),      # step = 2
]


def main():
    mg.import_reference("ViT-B-32")
    import src.services  # noqa: F401
    from src.services.frame_extractor import FrameExtractor

    out = []
    tmp = tempfile.mkdtemp(prefix="b200clip_frames_")
    for n, w, h, fps, rate in CASES:
        path = write_test_video(os.path.join(tmp, f"v_{n}_{w}_{rate}.mp4"), n, w, h, fps)
        fx = FrameExtractor()
        fx.sample_rate = rate
        frames, stamps = fx._extract_frames_opencv(path) if hasattr(fx, "_extract_frames_opencv") else fx.extract_frames(path)
        windows, wts = fx.create_sliding_windows(frames, stamps)
        out.append({"n": n, "w": w, "h": h, "fps": fps, "sample_rate": rate, "shape": list(frames.shape), "dtype": str(frames.dtype),
                    "timestamps": [float(t) for t in stamps], "window_timestamps": [float(t) for t in wts],
                    "windows": int(len(wts)), "frame_crc": [zlib.crc32(f.tobytes()) for f in frames]})
        print(n, w, h, fps, rate, "->", frames.shape, len(wts), "windows")
    with open(os.path.join(HERE, "frame_extractor.json"), "w") as f:
        json.dump({"cases": out}, f)


if __name__ == "__main__":
    main()
