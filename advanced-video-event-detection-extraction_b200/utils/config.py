"""Settings singleton with the reference's field names and defaults for the hot path
(/root/reference/src/utils/config.py:5-40,200-201).  Environment variables override the defaults, as with the
reference's pydantic BaseSettings (no .env parsing: only the fields this path reads are kept)."""
from __future__ import annotations

import os
from pathlib import Path


def _env(name, default, cast):
    v = os.environ.get(name)
    if v is None:
        return default
    try:
        return cast(v)
    except ValueError:
        return default


class Settings:
    def __init__(self):
        self.PROJECT_ROOT = Path(__file__).resolve().parent.parent.parent
        self.DATA_DIR = Path(_env("DATA_DIR", str(self.PROJECT_ROOT / "data"), str))
        self.MAX_VIDEO_SIZE = _env("MAX_VIDEO_SIZE", 2 * 1024 * 1024 * 1024, int)
        self.SUPPORTED_FORMATS = ["mp4", "avi", "mov", "mkv"]
        self.FRAME_SAMPLE_RATE = _env("FRAME_SAMPLE_RATE", 1, int)
        self.WINDOW_SIZE = _env("WINDOW_SIZE", 16, int)
        self.WINDOW_STRIDE = _env("WINDOW_STRIDE", 8, int)
        self.MAX_FRAME_WIDTH = _env("MAX_FRAME_WIDTH", 512, int)
        self.MAX_FRAME_HEIGHT = _env("MAX_FRAME_HEIGHT", 512, int)
        self.MAX_WINDOWS_PER_BATCH = _env("MAX_WINDOWS_PER_BATCH", 32, int)
        self.OPENCLIP_MODEL = _env("OPENCLIP_MODEL", "ViT-B-32", str)
        self.OPENCLIP_PRETRAINED = _env("OPENCLIP_PRETRAINED", "openai", str)
        self.BATCH_SIZE = _env("BATCH_SIZE", 32, int)
        self.TOP_K_RESULTS = _env("TOP_K_RESULTS", 15, int)
        self.CONFIDENCE_THRESHOLD = _env("CONFIDENCE_THRESHOLD", 0.25, float)
        self.CLIP_DURATION = _env("CLIP_DURATION", 30, int)
        self.MEMORY_CLEANUP_INTERVAL = _env("MEMORY_CLEANUP_INTERVAL", 5, int)
        # b200clip additions (not in the reference): frames per pass of the tower, frame cap compatibility
        self.B200_MAX_IMAGES_PER_PASS = _env("B200_MAX_IMAGES_PER_PASS", 1024, int)
        self.MAX_SAMPLED_FRAMES = _env("MAX_SAMPLED_FRAMES", 1000, int)  # frame_extractor.py:69-74
        # frame feed: decode only the frame each sliding window embeds (its middle frame) instead of every sampled frame
        # (FrameExtractor.extract_window_middles; same frames, timestamps and results, 1/8 of the decode work and host
        # memory at the default 16 / 8 windows); 0 = decode every sampled frame like frame_extractor.py:76-104
        self.B200_DECODE_MIDDLES_ONLY = bool(_env("B200_DECODE_MIDDLES_ONLY", 1, int))
        # decoder handles (threads) the middles-only decode spreads its seeks over (OpenCV releases the GIL while decoding)
        self.B200_DECODE_WORKERS = _env("B200_DECODE_WORKERS", min(8, os.cpu_count() or 1), int)
        # embed each video once into DATA_DIR/embeddings/*.b2emb and answer later queries from the cache (off = the
        # reference's behaviour: decode + embed on every query)
        self.B200_EMBEDDING_CACHE = bool(_env("B200_EMBEDDING_CACHE", 0, int))
        # segment merge of the phase-1 hits (phase3_advanced.py:29-81 temporal consistency): 0 = off (reference phase-1
        # result), 1 = on hits (timestamp +- 2.5 s), "clips" = on the clip intervals
        _m = _env("B200_TEMPORAL_MERGE", "0", str)
        self.B200_TEMPORAL_MERGE = "clips" if _m == "clips" else bool(_env("B200_TEMPORAL_MERGE", 0, int))


settings = Settings()
