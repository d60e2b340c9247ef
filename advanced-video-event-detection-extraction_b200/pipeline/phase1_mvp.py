"""Phase1MVP: drop-in for /root/reference/src/pipeline/phase1_mvp.py::Phase1MVP (:14-163).

Same call (`process_video(video_path, query, top_k=None, debug_mode=None)`), same result dicts
({'timestamp','confidence','phase':'phase1_mvp','window_index'}, descending confidence, thresholded at
settings.CONFIDENCE_THRESHOLD), `(results, debug_info)` in debug mode, ValueError when nothing could be processed.
What changes is how it is computed: instead of a Python loop embedding one middle frame at a time with a device
sync per frame (:74-121), all middle frames go through ONE batched K1->K3 pass and K4 does similarity + top-k +
threshold on the device.  The reference's memory-pressure branches (:66-71,111-118) have no equivalent here.
With torch.distributed initialised the middle frames shard across ranks (see ..distributed)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from .. import capi
from ..distributed import allgather_candidates, shard_range, world_info
from ..models.openclip_model import OpenCLIPModel
from ..services.frame_extractor import FrameExtractor
from ..utils.config import settings
from ..utils.logger import get_logger

logger = get_logger(__name__)


class Phase1MVP:
    def __init__(self, debug_mode: bool = False, clip_model: OpenCLIPModel | None = None):
        self.clip_model = clip_model if clip_model is not None else OpenCLIPModel()
        self.frame_extractor = FrameExtractor()
        self.debug_mode = debug_mode
        self._caches = {}
        self._fp = None

    # -------------------------------------------------------------------------------------------
    def process_video(self, video_path: str, query: str, top_k: int = None, debug_mode: bool = None):
        use_cache = settings.B200_EMBEDDING_CACHE and not (self.debug_mode if debug_mode is None else debug_mode)
        if use_cache:
            return self._process_video_cached(video_path, query, top_k)
        frames, timestamps = self.frame_extractor.extract_frames(video_path)
        return self.process_frames(frames, timestamps, query, top_k, debug_mode)

    def _process_video_cached(self, video_path: str, query: str, top_k: int = None):
        """Opt-in (settings.B200_EMBEDDING_CACHE): embed a video once into data/embeddings/*.b2emb -- the directory
        the reference reserves but never uses (README.md:208) -- and answer this and every later query from the
        cached window embeddings with K4 alone.  Same result dicts as the uncached path."""
        import os

        from ..services.embedding_cache import EmbeddingCache, cache_path_for

        cache_dir = str(settings.DATA_DIR / "embeddings")
        path = cache_path_for(video_path, cache_dir, settings.OPENCLIP_MODEL, self._fingerprint())
        cache = self._caches.get(path)
        if cache is None:
            if os.path.exists(path):
                cache = EmbeddingCache.load(self.clip_model, path)
            if cache is None or len(cache) != cache.meta.get("windows", len(cache)):     # absent or interrupted build
                frames, timestamps = self.frame_extractor.extract_frames(video_path)
                cache = EmbeddingCache.build(self.clip_model, frames, timestamps, dtype="float32", path=path,
                                             resume=os.path.exists(path))
            self._caches[path] = cache
        res = cache.query(query, top_k)
        for r in res:           # the reference's dicts carry no interval
            r.pop("start", None)
            r.pop("end", None)
        return res

    def _fingerprint(self) -> str:
        if self._fp is None:
            from ..services.embedding_cache import weights_fingerprint

            sd = getattr(self.clip_model, "_state_dict", None)
            self._fp = weights_fingerprint(sd) if sd else f"seed{getattr(self.clip_model, '_seed', 0)}"
        return self._fp

    def process_frames(self, frames: np.ndarray, timestamps: Sequence[float], query: str, top_k: int = None,
                       debug_mode: bool = None, video_duration: float = 0.0, return_device: bool = False):
        """The body of process_video after decode (phase1_mvp.py:36-163)."""
        if top_k is None:
            top_k = settings.TOP_K_RESULTS
        if debug_mode is not None:
            self.debug_mode = debug_mode
        mid_idx, window_ts = self.frame_extractor.window_middles(len(frames), list(timestamps))
        logger.info(f"Extracted {len(frames)} frames, created {len(mid_idx)} sliding windows")
        if not mid_idx:
            raise ValueError("No windows could be processed due to memory constraints")
        model = self.clip_model.model
        text_embedding = torch.from_numpy(self.clip_model.encode_text(query)).to(model.device)

        rank, world = world_info()
        m = len(mid_idx)
        lo, hi = shard_range(m, rank, world)
        k_eff = min(int(top_k), capi_max_k())
        ts_dev = torch.tensor(window_ts, dtype=torch.float64, device=model.device)
        if hi > lo:
            middle = np.ascontiguousarray(np.asarray(frames)[np.asarray(mid_idx[lo:hi])])
            if middle.dtype != np.uint8:
                middle = (middle * 255).astype(np.uint8)  # openclip_model.py:188-189
            emb = torch.empty(hi - lo, model.embed_dim, device=model.device, dtype=torch.float32)
            model.encode_frames_u8_host(middle, resize_mode=capi.RESIZE_REFERENCE, normalize=True, out=emb)
        else:
            emb = torch.empty(0, model.embed_dim, device=model.device, dtype=torch.float32)
        scores, idx, iv, cnt = model.sim_topk(emb, text_embedding, k_eff, settings.CONFIDENCE_THRESHOLD, ts_dev,
                                              index_base=lo, clip_duration=settings.CLIP_DURATION,
                                              video_duration=video_duration)
        if world > 1:
            cs, ci = allgather_candidates(scores, idx)
            scores, idx, iv, cnt = model.topk_merge(cs, ci, settings.CONFIDENCE_THRESHOLD, ts_dev,
                                                    settings.CLIP_DURATION, video_duration)
        if return_device:
            return scores, idx, iv, cnt
        n_hits = int(cnt[0].item())
        s_host, i_host = scores[0, :n_hits].cpu().numpy(), idx[0, :n_hits].cpu().numpy()
        results: List[Dict] = []
        for s, i in zip(s_host, i_host):
            results.append({"timestamp": window_ts[int(i)], "confidence": float(s), "phase": "phase1_mvp",
                            "window_index": int(i)})
        logger.info(f"Phase 1 found {len(results)} candidate events")
        if self.debug_mode:
            sims = model.similarity(emb, text_embedding)[:, 0].cpu().numpy()
            norms = emb.norm(dim=-1).cpu().numpy()
            debug_info = [{"window_index": lo + j, "timestamp": window_ts[lo + j], "similarity": float(sims[j]),
                           "image_embedding_norm": float(norms[j]),
                           "frame_shape": tuple(np.asarray(frames[mid_idx[lo + j]]).shape)} for j in range(hi - lo)]
            return results, debug_info
        return results


def capi_max_k() -> int:
    return 32
