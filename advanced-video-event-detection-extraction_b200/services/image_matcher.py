"""Image-query CLIP scoring (BASELINE config 5).

Mirrors the CLIP part of /root/reference/src/services/image_matcher.py::ImageMatcher:
  `_compute_clip_similarity(reference_image, frame)`  (:254-272)  cosine of the two image embeddings,
  `_single_stage_matching(reference_image, frames, timestamps, top_k, similarity_threshold)`  (:980-1018).
The reference re-encodes the reference image for EVERY frame and embeds frames one at a time; here the reference
image is embedded once, all frames go through one batched K1->K3 pass (transform-only resize, like
OpenCLIPModel.encode_images) and the similarities are one K4 pass over the embedding matrix.  The multi-stage
histogram / ORB / SSIM matchers of the reference are CPU OpenCV code outside the CLIP path and are not reproduced.

Ordering: the reference sorts with Python's stable `list.sort(reverse=True)`, so equal confidences keep FRAME order
(lower index first) -- the opposite of phase 1's np.argsort()[::-1].  K4's device top-k uses the phase-1 rule, so
this path takes the dense similarities (n floats) and applies the stable sort on the host."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from .. import capi
from ..models.openclip_model import OpenCLIPModel
from ..utils.logger import get_logger

logger = get_logger(__name__)


class ImageMatcher:
    def __init__(self, clip_model: OpenCLIPModel | None = None):
        self.clip_model = clip_model          # image_matcher.py:128 loads it lazily
        self.thresholds = {"clip_similarity": 0.7}          # image_matcher.py thresholds['clip_similarity']

    def _lazy_load_clip_model(self):
        if self.clip_model is None:
            self.clip_model = OpenCLIPModel()

    def _embed(self, images: np.ndarray) -> torch.Tensor:
        images = np.asarray(images)
        if images.ndim == 3:
            images = images[None]
        if images.dtype != np.uint8:
            images = (images * 255).astype(np.uint8)         # openclip_model.py:167-168
        model = self.clip_model.model
        out = torch.empty(len(images), model.embed_dim, device=model.device, dtype=torch.float32)
        model.encode_frames_u8_host(np.ascontiguousarray(images), resize_mode=capi.RESIZE_BICUBIC, normalize=True, out=out)
        return out

    def clip_similarities(self, reference_image: np.ndarray, frames: np.ndarray) -> np.ndarray:
        """float32 [n]: cosine(reference embedding, frame embedding) for every frame, one pass."""
        self._lazy_load_clip_model()
        ref = self._embed(reference_image)
        if len(frames) == 0:
            return np.zeros(0, np.float32)
        emb = self._embed(frames)
        return self.clip_model.model.similarity(emb, ref)[:, 0].cpu().numpy()

    def _compute_clip_similarity(self, reference_image: np.ndarray, frame: np.ndarray) -> float:
        return float(self.clip_similarities(reference_image, np.asarray(frame)[None])[0])

    def stage2_clip_filter(self, reference_image: np.ndarray, candidates: List[Dict]) -> List[Dict]:
        """Stage 2 of `_multi_stage_matching` (image_matcher.py:407-415): every candidate of the hash pre-filter (dicts
        with a 'frame') gets its 'clip_similarity' and the list is cut at thresholds['clip_similarity'], order kept.
        The reference embeds the reference image and one candidate frame per call; here the reference image is embedded
        once and the candidate frames in one batched pass (frames of one video share a shape; mixed shapes fall back to
        one pass per shape).  Stages 1, 3 and 4 (perceptual hash, SSIM, ORB / histogram) are CPU OpenCV code outside the
        CLIP path and stay with the caller."""
        if not candidates:
            return []
        by_shape: Dict[tuple, List[int]] = {}
        for i, c in enumerate(candidates):
            by_shape.setdefault(tuple(np.asarray(c["frame"]).shape), []).append(i)
        for idx in by_shape.values():
            sims = self.clip_similarities(reference_image, np.stack([np.asarray(candidates[i]["frame"]) for i in idx]))
            for i, sim in zip(idx, sims):
                candidates[i]["clip_similarity"] = float(sim)
        kept = [c for c in candidates if c["clip_similarity"] >= self.thresholds["clip_similarity"]]
        logger.info(f"Stage 2 filtered to {len(kept)} candidates")
        return kept

    @staticmethod
    def rank_single_stage(similarities: Sequence[float], timestamps: Sequence[float], top_k: int,
                          similarity_threshold: float) -> List[Dict]:
        """image_matcher.py:997-1016 on precomputed similarities."""
        sims = np.asarray(similarities, dtype=np.float32)
        order = np.argsort(-sims, kind="stable")[:max(int(top_k), 0)]     # stable descending: ties keep frame order
        out = []
        for i in order:
            c = float(sims[i])
            if c >= similarity_threshold:
                out.append({"timestamp": timestamps[int(i)], "confidence": c, "clip_similarity": c,
                            "method": "single_stage_matching", "frame_index": int(i)})
        return out

    def _single_stage_matching(self, reference_image: np.ndarray, frames: np.ndarray, timestamps: List[float],
                               top_k: int, similarity_threshold: float) -> List[Dict]:
        logger.info("Starting single-stage image matching...")
        return self.rank_single_stage(self.clip_similarities(reference_image, frames), timestamps, top_k,
                                      similarity_threshold)
