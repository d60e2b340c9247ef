"""FrameExtractor: the window / timestamp semantics of /root/reference/src/services/frame_extractor.py that the
query path depends on (:66-104 sampling + timestamps, :237-273 sliding windows).  Decoding itself (Decord / OpenCV,
CPU codec libraries) is outside the accelerated path: `extract_frames` uses OpenCV when it is installed and returns
RAW decoded frames -- the <=512x512 INTER_AREA shrink the reference applies here
(memory_manager.py:299-322) is folded into the K1 kernel (resize mode REFERENCE), so no CPU resize runs."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from ..utils.config import settings
from ..utils.logger import get_logger

logger = get_logger(__name__)


class FrameExtractor:
    def __init__(self):
        self.sample_rate = settings.FRAME_SAMPLE_RATE
        self.window_size = settings.WINDOW_SIZE
        self.window_stride = settings.WINDOW_STRIDE

    def sample_indices(self, total_frames: int) -> List[int]:
        """frame_extractor.py:66-74: every `sample_rate`-th frame, capped at 1000 frames by uniform sub-sampling."""
        idx = list(range(0, total_frames, self.sample_rate))
        cap = settings.MAX_SAMPLED_FRAMES
        if len(idx) > cap:
            step = len(idx) // cap
            idx = idx[::step][:cap]
        return idx

    @staticmethod
    def _video_props(cap, cv2) -> Tuple[int, float]:
        """(frame count, fps) with the reference's sanity check of the container's fps (frame_extractor.py:134-151)."""
        total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        fps = cap.get(cv2.CAP_PROP_FPS)
        if fps <= 0 or fps > 1000:
            logger.warning(f"Invalid FPS from OpenCV: {fps}")
            try:
                duration_ms = cap.get(cv2.CAP_PROP_POS_MSEC)
                fps = total / (duration_ms / 1000.0) if duration_ms > 0 else 30.0
            except Exception:
                fps = 30.0
        return total, fps

    def extract_frames(self, video_path: str) -> Tuple[np.ndarray, List[float]]:
        """Decode + sample (frame_extractor.py:128-204, the OpenCV path); a frame's timestamp is the position the decoder
        reports after the seek divided by fps (:183,201), which is frame_index / fps on well-formed files.  Returns raw
        RGB frames."""
        try:
            import cv2
        except ImportError as e:  # pragma: no cover
            raise RuntimeError("video decoding needs OpenCV; pass frames to Phase1MVP.process_frames instead") from e
        cap = cv2.VideoCapture(video_path)
        if not cap.isOpened():
            raise ValueError(f"Cannot open video: {video_path}")
        total, fps = self._video_props(cap, cv2)
        frames, stamps = [], []
        for i in self.sample_indices(total):
            cap.set(cv2.CAP_PROP_POS_FRAMES, i)
            actual_pos = cap.get(cv2.CAP_PROP_POS_FRAMES)
            ok, frame = cap.read()
            if not ok or frame is None:
                continue
            frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
            stamps.append(float(actual_pos) / float(fps))
        cap.release()
        if not frames:
            raise ValueError(f"No frames extracted from video: {video_path}")
        return np.stack(frames), stamps

    def extract_window_middles(self, video_path: str, bgr: bool = False, only=None):
        """Frame feed for phase 1: decode ONLY the frames phase 1 embeds -- the middle frame of every
        sliding window (phase1_mvp.py:80) -- instead of every sampled frame (frame_extractor.py:76-104 decodes all of
        them and phase 1 then drops 7 of 8 at the default 16 / 8 windows).  Same decoder calls per kept index as
        `extract_frames` (absolute seek + read), so the frames and timestamps are the reference's.  Returns
        `(middle_frames [m,H,W,3] uint8 RGB, window_timestamps [m], n_sampled)`, or None when the shortcut cannot prove
        that it is equivalent: a middle frame or the LAST sampled frame fails to decode (a truncated file shortens the
        reference's frame list and with it the windows) -- the caller then takes the full `extract_frames` path.
        `bgr=True` returns the frames as the decoder delivers them (OpenCV's BGR order) and leaves the channel swap of
        frame_extractor.py:191 to K1 (capi.INPUT_BGR), which saves a pass over every decoded frame on the host.
        `only=[window indices]` decodes just those windows' frames (returned in that order, with their timestamps):
        what a re-ranker needs for its <= 2 * top_k candidates."""
        try:
            import cv2
        except ImportError as e:  # pragma: no cover
            raise RuntimeError("video decoding needs OpenCV; pass frames to Phase1MVP.process_frames instead") from e
        cap = cv2.VideoCapture(video_path)
        if not cap.isOpened():
            raise ValueError(f"Cannot open video: {video_path}")
        try:
            total, fps = self._video_props(cap, cv2)
        finally:
            cap.release()
        sampled = self.sample_indices(total)
        if not sampled:
            return None
        # (nominal stamps only drive the window arithmetic; a window's timestamp is its middle frame's, and that is taken
        # from the decoder position reported after the seek, like the reference's timestamps list)
        mid_idx, _ = self.window_middles(len(sampled), [0.0] * len(sampled))
        if only is not None:
            if any(w < 0 or w >= len(mid_idx) for w in only):
                raise IndexError(f"window index out of range (video has {len(mid_idx)} windows)")
            mid_idx = [mid_idx[w] for w in only]
        mid_set = set(mid_idx)
        need = sorted(mid_set | {len(sampled) - 1})               # + sentinel: the last sampled frame must decode

        def read_some(js):
            """One decoder handle per worker; the reference's calls (absolute seek + read) per index."""
            c = cv2.VideoCapture(video_path)
            out = {}
            try:
                if not c.isOpened():
                    return None
                for j in js:
                    c.set(cv2.CAP_PROP_POS_FRAMES, sampled[j])
                    actual_pos = c.get(cv2.CAP_PROP_POS_FRAMES)
                    ok, frame = c.read()
                    if not ok or frame is None:
                        return None
                    if j in mid_set:
                        out[j] = (frame if bgr else cv2.cvtColor(frame, cv2.COLOR_BGR2RGB), float(actual_pos) / float(fps))
                return out
            finally:
                c.release()

        # seeks dominate (a seek decodes forward from the previous key frame), and OpenCV releases the GIL inside
        # read(): contiguous runs of the needed indices go to B200_DECODE_WORKERS decoder handles in parallel
        workers = max(1, min(int(settings.B200_DECODE_WORKERS), len(need) // 4 or 1))
        if workers == 1:
            parts = [read_some(need)]
        else:
            from concurrent.futures import ThreadPoolExecutor

            step = -(-len(need) // workers)
            with ThreadPoolExecutor(max_workers=workers) as pool:
                parts = list(pool.map(read_some, [need[i:i + step] for i in range(0, len(need), step)]))
        if any(p is None for p in parts):
            return None
        got = {}
        for p in parts:
            got.update(p)
        if not mid_idx:
            return np.empty((0, 0, 0, 3), np.uint8), [], len(sampled)
        return np.stack([got[j][0] for j in mid_idx]), [got[j][1] for j in mid_idx], len(sampled)

    def window_middles(self, n_frames: int, timestamps: List[float]) -> Tuple[List[int], List[float]]:
        """Index of the frame Phase 1 embeds for every sliding window (`window[len(window)//2]`,
        phase1_mvp.py:80) and the window timestamp (frame_extractor.py:237-273) -- without materialising the
        [m,16,H,W,3] window copy the reference builds."""
        if n_frames != len(timestamps):
            raise ValueError(f"Frames and timestamps length mismatch: {n_frames} vs {len(timestamps)}")
        if n_frames < self.window_size:
            logger.warning(f"Not enough frames ({n_frames}) for window size ({self.window_size})")
            if n_frames > 0:
                return [n_frames // 2], [timestamps[len(timestamps) // 2]]
            return [], []
        idx, ts = [], []
        for i in range(0, n_frames - self.window_size + 1, self.window_stride):
            mid = i + self.window_size // 2
            if mid >= len(timestamps):
                mid = len(timestamps) - 1
            idx.append(i + self.window_size // 2)
            ts.append(timestamps[mid])
        return idx, ts

    def create_sliding_windows(self, frames: np.ndarray, timestamps: List[float]):
        """Same return value as the reference (windows array + timestamps), for callers that want it."""
        if len(frames) != len(timestamps):
            raise ValueError(f"Frames and timestamps length mismatch: {len(frames)} vs {len(timestamps)}")
        if len(frames) < self.window_size:
            if len(frames) > 0:
                return np.array([frames]), [timestamps[len(timestamps) // 2]]
            return np.array([]), []
        windows, stamps = [], []
        for i in range(0, len(frames) - self.window_size + 1, self.window_stride):
            windows.append(frames[i:i + self.window_size])
            mid = min(i + self.window_size // 2, len(timestamps) - 1)
            stamps.append(timestamps[mid])
        return np.array(windows), stamps
