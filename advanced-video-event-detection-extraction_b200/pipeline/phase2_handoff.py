"""Phase-2 hand-off: the part of /root/reference/src/pipeline/phase2_reranker.py::Phase2Reranker (:10-90) that sits ON
the CLIP query path -- phase 1 is asked for 2 * top_k candidates (20 when top_k is None),
the middle frame of each candidate window is captioned, caption / query similarity is blended with the CLIP score as
0.7 * clip + 0.3 * caption, the hits are re-sorted (Python's stable descending sort: equal scores keep phase-1 order)
and cut to top_k.  Same call, same result dicts ('phase': 'phase2_reranked', 'caption', 'clip_score', 'caption_score').

The captioning model itself (BLIP, src/models/blip_model.py) is another model family and stays outside this repository:
it is INJECTED (`caption_model` with the reference's two calls, `generate_caption(frame) -> str` and
`compute_text_similarity(caption, query) -> float`); without one the class fails loudly instead of inventing scores.
What changes against the reference is the frame access: it decodes the whole video a second time and rebuilds every
window to fetch <= 2 * top_k frames (:51-52); here only those frames are decoded
(FrameExtractor.extract_window_middles(only=...)), with the reference's full path as the fallback."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

from ..services.frame_extractor import FrameExtractor
from ..utils.config import settings
from ..utils.logger import get_logger
from .phase1_mvp import Phase1MVP

logger = get_logger(__name__)

CLIP_WEIGHT, CAPTION_WEIGHT = 0.7, 0.3      # phase2_reranker.py:70


class Phase2Reranker:
    def __init__(self, phase1: Optional[Phase1MVP] = None, caption_model=None, frame_extractor=None):
        self.phase1 = phase1 if phase1 is not None else Phase1MVP()
        self.frame_extractor = frame_extractor if frame_extractor is not None else FrameExtractor()
        self.blip_model = caption_model

    def _ensure_blip_loaded(self):
        if self.blip_model is None:
            raise RuntimeError("Phase2Reranker needs a captioning model (generate_caption / compute_text_similarity); "
                               "BLIP is outside the accelerated path -- pass caption_model=")

    def _middle_frames(self, video_path: str, window_indices: Sequence[int]) -> Dict[int, "object"]:
        """window index -> the frame the reference captions (`window[len(window) // 2]`, :60-61)."""
        fx = self.frame_extractor
        wanted = sorted(set(int(w) for w in window_indices))
        if settings.B200_DECODE_MIDDLES_ONLY and hasattr(fx, "extract_window_middles") and "extract_frames" not in vars(fx):
            got = fx.extract_window_middles(video_path, only=wanted)
            if got is not None:
                return dict(zip(wanted, got[0]))
        frames, timestamps = fx.extract_frames(video_path)
        windows, _ = fx.create_sliding_windows(frames, timestamps)
        return {w: windows[w][len(windows[w]) // 2] for w in wanted}

    def process_video(self, video_path: str, query: str, top_k: int = None, debug_mode: bool = False) -> List[Dict]:
        logger.info(f"Phase 2 processing: {video_path} with query: '{query}'")
        self._ensure_blip_loaded()
        phase1_result = self.phase1.process_video(video_path, query, top_k * 2 if top_k else 20, debug_mode=debug_mode)
        phase1_results = phase1_result[0] if debug_mode and isinstance(phase1_result, tuple) else phase1_result
        if not phase1_results:
            return []
        middles = self._middle_frames(video_path, [r["window_index"] for r in phase1_results])
        reranked = []
        for result in phase1_results:
            window_idx = result["window_index"]
            caption = self.blip_model.generate_caption(middles[int(window_idx)])
            caption_similarity = self.blip_model.compute_text_similarity(caption, query)
            clip_score = result["confidence"]
            combined = CLIP_WEIGHT * clip_score + CAPTION_WEIGHT * caption_similarity
            reranked.append({"timestamp": result["timestamp"], "confidence": float(combined), "phase": "phase2_reranked",
                             "window_index": window_idx, "caption": caption, "clip_score": float(clip_score),
                             "caption_score": float(caption_similarity)})
        reranked.sort(key=lambda x: x["confidence"], reverse=True)
        final = reranked[:top_k or len(reranked)]
        logger.info(f"Phase 2 re-ranked to {len(final)} results")
        return final
