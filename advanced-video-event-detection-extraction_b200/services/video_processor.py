"""VideoProcessor.process_query (mode="mvp"): drop-in for the boundary caller at
/root/reference/src/services/video_processor.py:387-517 (+ preprocess_query :336-385, validate_video :817-847).
Only the text-query path is kept; the other model stacks the reference loads here are out of scope."""
from __future__ import annotations

import os
import re
from typing import Dict, Optional

from ..pipeline.phase1_mvp import Phase1MVP
from ..services.clip_extractor import ClipExtractor
from ..utils.config import settings
from ..utils.logger import get_logger

logger = get_logger(__name__)

_QUERY_IMPROVEMENTS = {
    r"\bwalks?\b": "walking", r"\bruns?\b": "running", r"\bjumps?\b": "jumping", r"\bfalls?\b": "falling",
    r"\bsits?\b": "sitting", r"\bstands?\b": "standing", r"\bdrives?\b": "driving", r"\bhits?\b": "hitting",
    r"\bcrashes?\b": "crashing",
    r"\bautomobile\b": "car", r"\bvehicle\b": "car", r"\bpedestrian\b": "person", r"\bindividual\b": "person",
    r"\bcanine\b": "dog",
    r"\bdark blue\b": "navy", r"\blight blue\b": "blue", r"\bdark green\b": "green", r"\blight green\b": "green",
}
_FILLERS = ["very", "really", "quite", "somewhat", "rather", "pretty"]


class VideoProcessor:
    def __init__(self, phase1: Phase1MVP | None = None, phase2=None):
        """`phase2`: a pipeline.phase2_handoff.Phase2Reranker built around an injected captioning model (BLIP itself is
        outside this path); without one "reranked" / "advanced" fall back to phase 1 like the reference does when its
        phase 2 is unavailable (video_processor.py:436-447)."""
        self.phase1 = phase1 if phase1 is not None else Phase1MVP()
        self.clip_extractor = ClipExtractor()
        self.phase2 = phase2
        self.phase2_available = phase2 is not None

    def preprocess_query(self, query: str) -> str:
        query = re.sub(r"\s+", " ", query.strip())
        query = query.lower()
        for pattern, replacement in _QUERY_IMPROVEMENTS.items():
            query = re.sub(pattern, replacement, query)
        query = re.sub(r"\b(a|an|the)\s+", "", query)
        for word in _FILLERS:
            query = re.sub(rf"\b{word}\s+", "", query)
        logger.info(f"Query preprocessed: '{query}'")
        return query

    def validate_video(self, video_path: str) -> Dict:
        if not os.path.exists(video_path):
            return {"valid": False, "error": f"Video file not found: {video_path}"}
        ext = os.path.splitext(video_path)[1].lower().lstrip(".")
        if ext not in settings.SUPPORTED_FORMATS:
            return {"valid": False, "error": f"Unsupported video format: {ext}"}
        if os.path.getsize(video_path) > settings.MAX_VIDEO_SIZE:
            return {"valid": False, "error": "Video file too large"}
        return {"valid": True}

    def process_query(self, video_path: str, query: str, mode: str = "mvp", top_k: Optional[int] = None,
                      threshold: Optional[float] = None, debug_mode: bool = False, merge=None) -> Dict:
        """`merge`: see Phase1MVP.process_video (None = settings.B200_TEMPORAL_MERGE, default off)."""
        if top_k is None:
            top_k = settings.TOP_K_RESULTS
        if threshold is None:
            threshold = settings.CONFIDENCE_THRESHOLD
        original_query = query
        processed_query = self.preprocess_query(query)
        try:
            validation = self.validate_video(video_path)
            if not validation["valid"]:
                return {"status": "error", "error": f"Video validation failed: {validation['error']}",
                        "query": original_query, "mode": mode, "results": []}
            if mode not in ("mvp", "reranked", "advanced"):
                raise ValueError(f"Unknown processing mode: {mode}")
            if mode != "mvp" and self.phase2_available:
                result = self.phase2.process_video(video_path, processed_query, top_k, debug_mode=debug_mode)
            else:
                if mode != "mvp":
                    logger.warning("Phase 2 not available, falling back to MVP mode")
                result = self.phase1.process_video(video_path, processed_query, top_k, debug_mode=debug_mode, merge=merge)
            debug_info = None
            if debug_mode and isinstance(result, tuple):
                results, debug_info = result
            else:
                results = result
            filtered = [r for r in results if isinstance(r, dict) and "confidence" in r and "timestamp" in r
                        and r["confidence"] >= threshold]
            for r in filtered:
                start, end = self.clip_extractor.clip_interval(r["timestamp"], settings.CLIP_DURATION)
                r["clip_path"] = None          # no ffmpeg transcode on this path
                r["clip_start"], r["clip_end"] = start, end
            response = {"status": "success", "query": original_query, "processed_query": processed_query,
                        "mode": mode, "results": filtered, "total_found": len(filtered)}
            if debug_info is not None:
                response["debug_info"] = debug_info
            return response
        except MemoryError as e:
            return {"status": "error", "error": f"Insufficient memory to process video. Details: {e}",
                    "query": original_query, "mode": mode, "results": [], "error_type": "memory_error"}
