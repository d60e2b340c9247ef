"""VideoProcessor.process_query (mode="mvp"): drop-in for the boundary caller at
/root/reference/src/services/video_processor.py:387-517 (+ preprocess_query :336-385, validate_video :817-847).
Only the text-query path is kept; the other model stacks the reference loads here are out of scope."""
from __future__ import annotations

import re
from pathlib import Path
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
        self.phase1 = phase1
        self.clip_extractor = ClipExtractor()
        self.phase2 = phase2
        self.phase2_available = phase2 is not None
        self._models_loaded = phase1 is not None
        if not self._models_loaded:
            # like the reference's constructor (video_processor.py:19-32,145-170): a failed load is logged, the object
            # survives, and process_query reports it as a 'model_loading_error' dict (:390-402) after one more attempt
            try:
                self._load_models()
            except Exception as e:
                logger.error(f"Model loading failed: {e}")

    def _load_models(self):
        self.phase1 = Phase1MVP()
        self._models_loaded = True

    def _ensure_models_loaded(self):
        if not self._models_loaded:
            self._load_models()

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
        """video_processor.py:817-847, message for message (they travel to the caller inside process_query's error
        dict)."""
        try:
            video_file = Path(video_path)
            if not video_file.exists():
                return {"valid": False, "error": "Video file does not exist"}
            file_extension = video_file.suffix.lower().lstrip(".")
            if file_extension not in settings.SUPPORTED_FORMATS:
                return {"valid": False,
                        "error": f"Unsupported format: {file_extension}. Supported: {settings.SUPPORTED_FORMATS}"}
            file_size = video_file.stat().st_size
            if file_size > settings.MAX_VIDEO_SIZE:
                return {"valid": False,
                        "error": f"Video file too large: {file_size} bytes. Max: {settings.MAX_VIDEO_SIZE} bytes"}
            return {"valid": True, "format": file_extension, "size": file_size, "path": str(video_file)}
        except Exception as e:
            return {"valid": False, "error": f"Error validating video: {str(e)}"}

    def process_query(self, video_path: str, query: str, mode: str = "mvp", top_k: Optional[int] = None,
                      threshold: Optional[float] = None, debug_mode: bool = False, merge=None) -> Dict:
        """`merge`: see Phase1MVP.process_video (None = settings.B200_TEMPORAL_MERGE, default off)."""
        try:
            self._ensure_models_loaded()
        except Exception as e:
            return {"status": "error", "error": f"Failed to load required models: {str(e)}", "query": query, "mode": mode,
                    "results": [], "error_type": "model_loading_error"}
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
            return {"status": "error",
                    "error": "Insufficient memory to process video. Try using a smaller video or restart the application. "
                             f"Details: {str(e)}",
                    "query": original_query, "mode": mode, "results": [], "error_type": "memory_error"}
