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
from ..distributed import shard_range, world_info
from ..models.openclip_model import OpenCLIPModel
from ..services.frame_extractor import FrameExtractor
from ..utils.config import settings
from ..utils.logger import get_logger

logger = get_logger(__name__)


def shrunk_shape(shape) -> tuple:
    """Shape of a decoded frame after the reference's <= MAX_FRAME_WIDTH x MAX_FRAME_HEIGHT shrink
    (memory_manager.py:299-312: scale = min(max_w / w, max_h / h, 1), int() truncation) -- the frame phase 1 embeds
    there and reports as `frame_shape` in debug mode; here the shrink happens inside K1."""
    h, w = int(shape[0]), int(shape[1])
    scale = min(settings.MAX_FRAME_WIDTH / w, settings.MAX_FRAME_HEIGHT / h, 1.0)
    if scale < 1.0:
        return (int(h * scale), int(w * scale)) + tuple(shape[2:])
    return tuple(shape)


def dump_debug_frames(debug_dir, frames_rgb, window_indices, similarities, n_windows: int) -> list:
    """phase1_mvp.py:99-102: in debug mode the embedded frame of the first and last five windows is written to
    DATA_DIR/debug/frame_<i>_sim_<s>.jpg -- the frame AFTER the <= 512 x 512 INTER_AREA shrink, as the reference holds
    it.  A debugging artefact on the host (at most ten small frames), not part of the query path."""
    try:
        import cv2
    except ImportError:  # pragma: no cover
        return []
    from pathlib import Path

    debug_dir = Path(debug_dir)
    debug_dir.mkdir(parents=True, exist_ok=True)
    written = []
    for frame, i, sim in zip(frames_rgb, window_indices, similarities):
        if not (i < 5 or i >= n_windows - 5):
            continue
        frame = np.asarray(frame)
        hh, ww = shrunk_shape(frame.shape)[:2]
        if (hh, ww) != frame.shape[:2]:
            frame = cv2.resize(frame, (ww, hh), interpolation=cv2.INTER_AREA)
        path = debug_dir / f"frame_{int(i):03d}_sim_{float(sim):.4f}.jpg"
        cv2.imwrite(str(path), cv2.cvtColor(frame, cv2.COLOR_RGB2BGR))
        written.append(str(path))
    return written


class Phase1MVP:
    def __init__(self, debug_mode: bool = False, clip_model: OpenCLIPModel | None = None):
        self.clip_model = clip_model if clip_model is not None else OpenCLIPModel()
        self.frame_extractor = FrameExtractor()
        self.debug_mode = debug_mode
        self._caches = {}
        self._fp = None

    # -------------------------------------------------------------------------------------------
    def process_video(self, video_path: str, query: str, top_k: int = None, debug_mode: bool = None, merge=None):
        """`merge` (b200clip addition, default settings.B200_TEMPORAL_MERGE = off -> the reference's phase-1 result):
        True  -> Phase3Advanced._apply_temporal_consistency + confidence sort on the hits (timestamp +- 2.5 s segments,
                 /root/reference/src/pipeline/phase3_advanced.py:29-81);
        "clips" -> the same filter on the K4 clip intervals (start_time/end_time = clip_extractor.py:175-183), i.e. the
                 hits that survive are clips no two of which overlap by more than half."""
        use_cache = settings.B200_EMBEDDING_CACHE and not (self.debug_mode if debug_mode is None else debug_mode)
        if use_cache and world_info()[1] > 1:
            # every rank would append to the same file, and the cached path does not shard: one writer only
            logger.warning("B200_EMBEDDING_CACHE is ignored while torch.distributed is initialised (world > 1)")
            use_cache = False
        if use_cache:
            return self._merge(self._process_video_cached(video_path, query, top_k, keep_intervals=True), merge)
        fx = self.frame_extractor
        if settings.B200_DECODE_MIDDLES_ONLY and "extract_frames" not in vars(fx):     # a replaced decoder stays in charge
            got = fx.extract_window_middles(video_path, bgr=True)     # BGR2RGB (frame_extractor.py:191) happens in K1
            if got is not None:
                middle, window_ts, n_sampled = got
                logger.info(f"Extracted {len(middle)} window-middle frames of {n_sampled} sampled frames")
                return self.process_middle_frames(middle, window_ts, query, top_k, debug_mode, merge=merge, bgr=True)
        frames, timestamps = fx.extract_frames(video_path)
        return self.process_frames(frames, timestamps, query, top_k, debug_mode, merge=merge)

    @staticmethod
    def _merge(results, merge):
        """Segment merge of a hit list (see process_video); always strips the helper interval keys of unmerged hits."""
        if merge is None:
            merge = settings.B200_TEMPORAL_MERGE
        if merge:
            from .temporal import merge_hits

            if merge == "clips":
                hits = [dict(r, start_time=r["start"], end_time=r["end"]) for r in results]
            else:
                hits = [{k: v for k, v in r.items() if k not in ("start", "end")} for r in results]
            return [{k: v for k, v in r.items() if k not in ("start", "end")} for r in merge_hits(hits)]
        return [{k: v for k, v in r.items() if k not in ("start", "end")} for r in results]

    def _process_video_cached(self, video_path: str, query: str, top_k: int = None, keep_intervals: bool = False):
        """Opt-in (settings.B200_EMBEDDING_CACHE): embed a video once into data/embeddings/*.b2emb -- the directory
        the reference reserves but never uses (README.md:208) -- and answer this and every later query from the
        cached window embeddings with K4 alone.  Same result dicts as the uncached path."""
        import os

        from ..services.embedding_cache import EmbeddingCache, cache_path_for

        cache_dir = str(settings.DATA_DIR / "embeddings")
        expect = {"weights_fingerprint": self._fingerprint(), "window_size": settings.WINDOW_SIZE,
                  "window_stride": settings.WINDOW_STRIDE, "frame_sample_rate": settings.FRAME_SAMPLE_RATE,
                  "max_sampled_frames": settings.MAX_SAMPLED_FRAMES, "resize_mode": int(capi.RESIZE_REFERENCE),
                  "model": settings.OPENCLIP_MODEL, "pretrained": settings.OPENCLIP_PRETRAINED}
        sampling = "|".join(f"{k}={expect[k]}" for k in sorted(expect) if k != "weights_fingerprint")
        path = cache_path_for(video_path, cache_dir, settings.OPENCLIP_MODEL, self._fingerprint(), sampling)
        cache = self._caches.get(path)
        if cache is None:
            resume = os.path.exists(path)
            if resume:
                try:
                    cache = EmbeddingCache.load(self.clip_model, path, expect=expect)
                except (ValueError, OSError, KeyError) as e:       # torn header / other settings: treat as absent
                    logger.warning(f"embedding cache {path} is unusable ({e}); rebuilding")
                    cache, resume = None, False
                    os.replace(path, path + ".bad")
            if cache is None or len(cache) != cache.meta.get("windows", len(cache)):     # absent or interrupted build
                fx = self.frame_extractor
                got = None
                if settings.B200_DECODE_MIDDLES_ONLY and "extract_frames" not in vars(fx):
                    got = fx.extract_window_middles(video_path, bgr=True)
                if got is not None:
                    cache = EmbeddingCache.build_from_middles(self.clip_model, got[0], got[1], dtype="float32", path=path,
                                                              resume=resume, fingerprint=self._fingerprint(),
                                                              resize_mode=capi.RESIZE_REFERENCE | capi.INPUT_BGR)
                else:
                    frames, timestamps = fx.extract_frames(video_path)
                    cache = EmbeddingCache.build(self.clip_model, frames, timestamps, dtype="float32", path=path,
                                                 resume=resume, fingerprint=self._fingerprint())
            self._caches[path] = cache
        res = cache.query(query, top_k)
        if not keep_intervals:
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
                       debug_mode: bool = None, video_duration: float = 0.0, return_device: bool = False, merge=None):
        """The body of process_video after decode (phase1_mvp.py:36-163)."""
        mid_idx, window_ts = self.frame_extractor.window_middles(len(frames), list(timestamps))
        logger.info(f"Extracted {len(frames)} frames, created {len(mid_idx)} sliding windows")
        fr = np.asarray(frames)
        return self._process_middles(lambda lo, hi: fr[np.asarray(mid_idx[lo:hi], dtype=np.int64)], window_ts, query, top_k,
                                     debug_mode, video_duration, return_device, merge)

    def process_middle_frames(self, middle_frames: np.ndarray, window_timestamps: Sequence[float], query: str,
                              top_k: int = None, debug_mode: bool = None, video_duration: float = 0.0,
                              return_device: bool = False, merge=None, bgr: bool = False):
        """process_frames for a caller that already holds ONLY the frame each window embeds (its middle frame,
        phase1_mvp.py:80) and the window timestamps -- what FrameExtractor.extract_window_middles decodes.
        `bgr=True`: the frames are in OpenCV's BGR order (K1 swaps the channels in its final store)."""
        if len(middle_frames) != len(window_timestamps):
            raise ValueError(f"Frames and timestamps length mismatch: {len(middle_frames)} vs {len(window_timestamps)}")
        fr = np.asarray(middle_frames)
        return self._process_middles(lambda lo, hi: fr[lo:hi], list(window_timestamps), query, top_k, debug_mode,
                                     video_duration, return_device, merge, bgr)

    def _process_middles(self, middles_of, window_ts, query, top_k, debug_mode, video_duration, return_device, merge,
                         bgr: bool = False):
        """`middles_of(lo, hi)` -> the frames windows [lo, hi) embed, as one array."""
        if top_k is None:
            top_k = settings.TOP_K_RESULTS
        if debug_mode is not None:
            self.debug_mode = debug_mode
        if not window_ts:
            raise ValueError("No windows could be processed due to memory constraints")
        model = self.clip_model.model
        text_embedding = torch.from_numpy(self.clip_model.encode_text(query)).to(model.device)

        rank, world = world_info()
        lo, hi = shard_range(len(window_ts), rank, world)
        k_eff = int(top_k)      # any top_k: K4 serves k > 32 in ceil(k / 32) passes (phase1_mvp.py:145 takes any value)
        if k_eff <= 0:
            return ([], []) if self.debug_mode else []
        ts_dev = torch.tensor(window_ts, dtype=torch.float64, device=model.device)
        if hi > lo:
            middle = np.ascontiguousarray(middles_of(lo, hi))
            if middle.dtype != np.uint8:
                middle = (middle * 255).astype(np.uint8)  # openclip_model.py:188-189
            emb = torch.empty(hi - lo, model.embed_dim, device=model.device, dtype=torch.float32)
            model.encode_frames_u8_host(middle, resize_mode=capi.RESIZE_REFERENCE | (capi.INPUT_BGR if bgr else 0),
                                        normalize=True, out=emb)
        else:
            emb = torch.empty(0, model.embed_dim, device=model.device, dtype=torch.float32)
        # (world > 1: local top-k -> one all-gather of the packed candidates -> merge, see ..distributed)
        scores, idx, iv, cnt = model.sim_topk_sharded(emb, text_embedding, k_eff, settings.CONFIDENCE_THRESHOLD, ts_dev,
                                                      index_base=lo, clip_duration=settings.CLIP_DURATION,
                                                      video_duration=video_duration)
        if return_device:
            return scores, idx, iv, cnt
        n_hits = int(cnt[0].item())
        s_host, i_host = scores[0, :n_hits].cpu().numpy(), idx[0, :n_hits].cpu().numpy()
        iv_host = iv[0, :n_hits].cpu().numpy()
        results: List[Dict] = []
        for s, i, (a, b) in zip(s_host, i_host, iv_host):
            results.append({"timestamp": window_ts[int(i)], "confidence": float(s), "phase": "phase1_mvp",
                            "window_index": int(i), "start": float(a), "end": float(b)})
        results = self._merge(results, merge)
        logger.info(f"Phase 1 found {len(results)} candidate events")
        if self.debug_mode:
            sims = model.similarity(emb, text_embedding)[:, 0].cpu().numpy()
            norms = emb.norm(dim=-1).cpu().numpy()
            debug_info = [{"window_index": lo + j, "timestamp": window_ts[lo + j], "similarity": float(sims[j]),
                           "image_embedding_norm": float(norms[j]),
                           "frame_shape": shrunk_shape(middle.shape[1:])} for j in range(hi - lo)]
            if len(sims):
                edge = [j for j in range(hi - lo) if lo + j < 5 or lo + j >= len(window_ts) - 5]
                dump_debug_frames(settings.DATA_DIR / "debug", [middle[j][..., ::-1] if bgr else middle[j] for j in edge],
                                  [lo + j for j in edge], [sims[j] for j in edge], len(window_ts))
                self._log_debug_analysis(sims, debug_info, query, settings.CONFIDENCE_THRESHOLD)
            return results, debug_info
        return results

    def _log_debug_analysis(self, similarities, debug_info, query, threshold) -> Dict:
        """phase1_mvp.py:165-212: score statistics, best / worst windows, share above the threshold and -- when nothing
        passes -- percentile-based threshold suggestions.  Logged like the reference; also returned (and kept in
        `self.last_debug_analysis`) so that a caller can read the numbers without parsing the log."""
        sims = np.asarray(similarities, dtype=np.float32)

        def ts_of(i):
            return debug_info[i].get("timestamp", 0.0) if i < len(debug_info) and isinstance(debug_info[i], dict) else 0.0

        top = [int(i) for i in np.argsort(sims)[::-1][:10]]
        bottom = [int(i) for i in np.argsort(sims)[:5]]
        above = int(np.sum(sims >= threshold))
        out = {"query": query, "windows": int(len(sims)), "min": float(sims.min()), "max": float(sims.max()),
               "mean": float(sims.mean()), "std": float(sims.std()), "threshold": threshold,
               "top10": [(i, float(sims[i]), ts_of(i)) for i in top],
               "bottom5": [(i, float(sims[i]), ts_of(i)) for i in bottom], "above_threshold": above, "suggested": []}
        logger.info("=== DEBUG ANALYSIS ===")
        logger.info(f"Query: '{query}'")
        logger.info(f"Total windows processed: {len(sims)}")
        logger.info(f"Similarity range: [{out['min']:.6f}, {out['max']:.6f}]")
        logger.info(f"Mean similarity: {out['mean']:.6f}")
        logger.info(f"Std similarity: {out['std']:.6f}")
        logger.info(f"Confidence threshold: {threshold}")
        logger.info("Top 10 similarity scores:")
        for n, (i, s, t) in enumerate(out["top10"]):
            logger.info(f"  {n + 1}. Window {i}: {s:.6f} at {t:.2f}s")
        logger.info("Bottom 5 similarity scores:")
        for n, (i, s, t) in enumerate(out["bottom5"]):
            logger.info(f"  {n + 1}. Window {i}: {s:.6f} at {t:.2f}s")
        logger.info(f"Windows above threshold ({threshold}): {above}/{len(sims)} ({above / len(sims) * 100:.1f}%)")
        if above == 0:
            logger.info("Suggested thresholds based on percentiles:")
            for p in (95, 90, 80, 70, 50):
                thresh = float(np.percentile(sims, p))
                count = int(np.sum(sims >= thresh))
                out["suggested"].append((p, thresh, count))
                logger.info(f"  {p}th percentile ({thresh:.4f}): {count} windows")
        logger.info("=== END DEBUG ANALYSIS ===")
        self.last_debug_analysis = out
        return out
