"""ORACLE (test infrastructure, not product): numpy / pure-Python restatement of the reference's
phase1_mvp query logic around the model: windows, similarity, top-k, thresholds, clip intervals, query text
normalisation.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Each function cites the reference file:line it follows (paths under /root/reference).  Pinned by
tests/golden/phase1_*.npz / .json, produced by running the reference's own unmodified classes
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import re

import numpy as np

WINDOW_SIZE = 16          # src/utils/config.py:15
WINDOW_STRIDE = 8         # src/utils/config.py:16
TOP_K_RESULTS = 15        # src/utils/config.py:38
CONFIDENCE_THRESHOLD = 0.25  # src/utils/config.py:39
CLIP_DURATION = 30        # src/utils/config.py:40
MAX_SAMPLED_FRAMES = 1000  # src/services/frame_extractor.py:69-74


def sliding_window_middles(n_frames: int, timestamps, window: int = WINDOW_SIZE, stride: int = WINDOW_STRIDE):
    """FrameExtractor.create_sliding_windows (src/services/frame_extractor.py:237-273) reduced to what
    Phase1MVP uses (src/pipeline/phase1_mvp.py:80: the middle frame `window[len(window)//2]`):
    returns (frame index embedded per window, window timestamp per window)."""
    if len(timestamps) != n_frames:
        raise ValueError(f"Frames and timestamps length mismatch: {n_frames} vs {len(timestamps)}")
    if n_frames < window:
        if n_frames > 0:
            # single window of all frames: embedded frame = frames[n//2], timestamp = timestamps[n//2]
            return [n_frames // 2], [timestamps[len(timestamps) // 2]]
        return [], []
    idx, ts = [], []
    for i in range(0, n_frames - window + 1, stride):
        mid = i + window // 2
        if mid >= len(timestamps):
            mid = len(timestamps) - 1
        idx.append(i + window // 2)
        ts.append(timestamps[mid])
    return idx, ts


def compute_similarity(image_embeddings: np.ndarray, text_embeddings: np.ndarray) -> np.ndarray:
    """OpenCLIPModel.compute_similarity (src/models/openclip_model.py:212-214)."""
    return np.dot(image_embeddings, text_embeddings.T)


def topk_threshold(similarities: np.ndarray, window_timestamps, top_k: int = TOP_K_RESULTS,
                   threshold: float = CONFIDENCE_THRESHOLD):
    """src/pipeline/phase1_mvp.py:134,145-155: argsort descending (ties -> higher index first), keep
    entries >= threshold, emit result dicts in that order."""
    similarities = np.asarray(similarities)
    top_indices = np.argsort(similarities)[::-1][:top_k]
    results = []
    for idx in top_indices:
        if similarities[idx] >= threshold:
            results.append({
                "timestamp": window_timestamps[idx],
                "confidence": float(similarities[idx]),
                "phase": "phase1_mvp",
                "window_index": int(idx),
            })
    return results


def topk_indices(similarities: np.ndarray, top_k: int) -> np.ndarray:
    return np.argsort(np.asarray(similarities))[::-1][:top_k]


def clip_interval(timestamp: float, duration: float = CLIP_DURATION, video_duration: float | None = None):
    """ClipExtractor.extract_clip_with_padding + the clamps at the top of extract_clip
    (src/services/clip_extractor.py:175-183, 94-111).  Returns (start, end) seconds."""
    start = max(0, timestamp - duration / 2)
    end = timestamp + duration / 2
    if start < 0:
        start = 0
    if end <= start:
        end = start + 5.0
    if video_duration:
        if start >= video_duration:
            start = max(0, video_duration - 5.0)
            end = video_duration
        elif end > video_duration:
            end = video_duration
    return float(start), float(end)


def second_threshold(results, threshold: float):
    """VideoProcessor.process_query filter (src/services/video_processor.py:463-471)."""
    return [r for r in results if isinstance(r, dict) and "confidence" in r and "timestamp" in r
            and r["confidence"] >= threshold]


_QUERY_IMPROVEMENTS = {
    r"\bwalks?\b": "walking", r"\bruns?\b": "running", r"\bjumps?\b": "jumping", r"\bfalls?\b": "falling",
    r"\bsits?\b": "sitting", r"\bstands?\b": "standing", r"\bdrives?\b": "driving", r"\bhits?\b": "hitting",
    r"\bcrashes?\b": "crashing",
    r"\bautomobile\b": "car", r"\bvehicle\b": "car", r"\bpedestrian\b": "person", r"\bindividual\b": "person",
    r"\bcanine\b": "dog",
    r"\bdark blue\b": "navy", r"\blight blue\b": "blue", r"\bdark green\b": "green", r"\blight green\b": "green",
}
_FILLERS = ["very", "really", "quite", "somewhat", "rather", "pretty"]


def preprocess_query(query: str) -> str:
    """VideoProcessor.preprocess_query (src/services/video_processor.py:336-385)."""
    query = re.sub(r"\s+", " ", query.strip())
    query = query.lower()
    for pattern, replacement in _QUERY_IMPROVEMENTS.items():
        query = re.sub(pattern, replacement, query)
    query = re.sub(r"\b(a|an|the)\s+", "", query)
    for word in _FILLERS:
        query = re.sub(rf"\b{word}\s+", "", query)
    return query


def merge_topk_lists(cand_scores: np.ndarray, cand_idx: np.ndarray, k: int):
    """What the multi-GPU merge must equal: the global argsort-descending top-k over the union of
    per-shard candidates (ties -> higher global index first).  cand_* are [g, k]; idx < 0 = empty."""
    s = cand_scores.reshape(-1)
    i = cand_idx.reshape(-1)
    keep = i >= 0
    s, i = s[keep], i[keep]
    order = np.lexsort((i, s))[::-1][:k]   # primary: score, secondary: index; descending both
    return s[order], i[order]


# ---------------------------------------------------------------------------- section 8f "next" rows
def temporal_consistency(results):
    """Phase3Advanced._apply_temporal_consistency, /root/reference/src/pipeline/phase3_advanced.py:37-81, restated
    statement by statement (including `filtered_results.remove(existing)` inside the loop over filtered_results,
    which makes Python's list iterator skip the element after a removal).  Pinned by tests/golden/next_rows.json."""
    if len(results) <= 1:
        return results
    sorted_results = sorted(results, key=lambda x: x["timestamp"])
    filtered = []
    for current in sorted_results:
        should_add = True
        for existing in filtered:                       # iterator over a list that may shrink underneath it
            c0 = current.get("start_time", current["timestamp"] - 2.5)
            c1 = current.get("end_time", current["timestamp"] + 2.5)
            e0 = existing.get("start_time", existing["timestamp"] - 2.5)
            e1 = existing.get("end_time", existing["timestamp"] + 2.5)
            overlap = max(0, min(c1, e1) - max(c0, e0))
            if overlap > 0.5 * (c1 - c0) or overlap > 0.5 * (e1 - e0):
                if current["confidence"] <= existing["confidence"]:
                    should_add = False
                    break
                else:
                    filtered.remove(existing)
        if should_add:
            filtered.append(current)
    return filtered


def single_stage_matching(similarities, timestamps, top_k, similarity_threshold):
    """ImageMatcher._single_stage_matching, /root/reference/src/services/image_matcher.py:980-1018, after the
    per-frame CLIP similarities: build the dicts, stable sort descending, cut to top_k, threshold."""
    sims = [{"timestamp": timestamps[i], "confidence": float(s), "clip_similarity": float(s),
             "method": "single_stage_matching", "frame_index": i} for i, s in enumerate(similarities)]
    sims.sort(key=lambda x: x["confidence"], reverse=True)
    return [s for s in sims[:top_k] if s["confidence"] >= similarity_threshold]
