"""Golden for the boundary caller's validation / error dicts (SURVEY.md section 8b "Error conventions"): the REFERENCE'S
OWN VideoProcessor.validate_video (/root/reference/src/services/video_processor.py:817-847) and the dicts
process_query returns for a failed validation and for a MemoryError raised by phase 1 (:417-425, :508-517), on files
created in a scratch directory (the directory is replaced by <DIR> in the stored strings).

  python tests/golden/make_golden_validate.py   (needs /root/reference; tests/golden/validate_video.json is committed)
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402

FILES = [("ok.mp4", 10), ("UPPER.MKV", 3), ("clip.avi", 0), ("movie.mov", 7), ("notes.txt", 4), ("noext", 4), ("big.mp4", 5000)]


def make_files(d):
    for name, size in FILES:
        with open(os.path.join(d, name), "wb") as f:
            f.write(b"x" * size)


def main():
    config = mg.import_reference("ViT-B-32")
    import src.services  # noqa: F401
    from src.services.video_processor import VideoProcessor

    d = tempfile.mkdtemp(prefix="b200clip_validate_")
    make_files(d)
    vp = object.__new__(VideoProcessor)          # no model loading: only the methods under test are used
    vp._models_loaded = True

    def scrub(x):
        return json.loads(json.dumps(x).replace(d, "<DIR>"))

    out = {"max_video_size": 4096, "validate": [], "process_query": []}
    config.settings.MAX_VIDEO_SIZE = 4096
    for name in [n for n, _ in FILES] + ["missing.mp4"]:
        out["validate"].append({"name": name, "result": scrub(VideoProcessor.validate_video(vp, os.path.join(d, name)))})
    for name in ("missing.mp4", "notes.txt", "big.mp4"):
        out["process_query"].append({"name": name, "query": "A  Dog jumps", "mode": "mvp",
                                     "result": scrub(VideoProcessor.process_query(vp, os.path.join(d, name), "A  Dog jumps"))})

    class P1:
        def process_video(self, *a, **k):
            raise MemoryError("cannot allocate 12 GB")

    vp.phase1 = P1()
    out["process_query"].append({"name": "ok.mp4", "query": "red car", "mode": "mvp", "memory_error": "cannot allocate 12 GB",
                                 "result": scrub(VideoProcessor.process_query(vp, os.path.join(d, "ok.mp4"), "red car"))})
    # a model load that fails again at query time (:390-402)
    vp._models_loaded = False
    vp._load_models = lambda: (_ for _ in ()).throw(RuntimeError("no checkpoint on this host"))
    out["process_query"].append({"name": "ok.mp4", "query": "red car", "mode": "reranked", "load_error": "no checkpoint on this host",
                                 "result": scrub(VideoProcessor.process_query(vp, os.path.join(d, "ok.mp4"), "red car", mode="reranked"))})
    with open(os.path.join(HERE, "validate_video.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote validate_video.json")


if __name__ == "__main__":
    main()
