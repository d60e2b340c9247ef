"""Golden for the clip interval: the REFERENCE'S OWN
ClipExtractor.extract_clip_with_padding -> extract_clip (/root/reference/src/services/clip_extractor.py:87-111,175-183),
with ffmpeg replaced by a recorder of the `ss` / `t` it is handed and the container probes (_validate_video_file,
_get_video_duration) by constants.  Stored: (timestamp, clip duration, video duration) -> (start, end) = (ss, ss + t).

  python tests/golden/make_golden_clips.py     (needs /root/reference; tests/golden/clip_intervals.json is committed)
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402

CALLS = []


class _Chain:
    def __init__(self, out=None):
        self.out = out

    def output(self, path, **_kw):
        return _Chain(path)

    def overwrite_output(self):
        return self

    def run(self, **_kw):
        with open(self.out, "wb") as f:
            f.write(b"clip")


def _input(video_path, ss=None, t=None, **_kw):
    CALLS.append((float(ss), float(t)))
    return _Chain()


def main():
    stub = types.ModuleType("ffmpeg")
    stub.input = _input
    stub.Error = type("Error", (Exception,), {"stderr": b""})
    stub.probe = lambda *_a, **_k: {"format": {"duration": "0"}}
    sys.modules["ffmpeg"] = stub
    mg.import_reference("ViT-B-32")
    import src.services  # noqa: F401
    from src.services.clip_extractor import ClipExtractor

    ce = ClipExtractor()
    ce._validate_video_file = lambda p: True
    rng = np.random.default_rng(5)
    cases = []
    grid = [(2.0, 30, None), (100.0, 30, 110.0), (100.0, 30, 90.0), (50.0, 0, None), (0.0, 30, 5.0), (14.99, 30, 1000.0),
            (300.0, 30, 100.0), (15.0, 30, 30.0), (15.0, None, 20.0), (3.0, 10, 3.0), (7.5, 5, 7.5), (0.0, 0, 0.0)]
    for _ in range(40):
        vd = float(rng.choice([0.0, 12.0, 61.5, 3600.0]))
        grid.append((round(float(rng.uniform(0, max(vd, 20.0) * 1.2)), 3), [None, 30, 10, 4.5][int(rng.integers(0, 4))],
                     vd if vd > 0 else None))
    for ts, dur, vd in grid:
        ce._get_video_duration = lambda p, vd=vd: vd
        CALLS.clear()
        ce.extract_clip_with_padding("video.mp4", ts, dur)
        (ss, t), = CALLS
        cases.append({"timestamp": ts, "duration": dur, "video_duration": vd, "start": ss, "end": ss + t, "t": t})
    with open(os.path.join(HERE, "clip_intervals.json"), "w") as f:
        json.dump({"cases": cases}, f, indent=0)
    print("wrote clip_intervals.json:", len(cases), "cases")


if __name__ == "__main__":
    main()
