"""Clip interval arithmetic of /root/reference/src/services/clip_extractor.py (:175-183 padding, :94-111 clamps).
The ffmpeg transcode the reference runs afterwards is an external binary and outside this path; the GPU emits
the same intervals from K4 (b200clip_sim_topk `intervals`), this module is the host-side spelling used when a
caller wants one interval."""
from __future__ import annotations

from typing import Optional, Tuple

from ..utils.config import settings


class ClipExtractor:
    def clip_interval(self, timestamp: float, duration: Optional[float] = None,
                      video_duration: Optional[float] = None) -> Tuple[float, float]:
        if duration is None:
            duration = settings.CLIP_DURATION
        start_time = max(0, timestamp - duration / 2)
        end_time = timestamp + duration / 2
        if start_time < 0:
            start_time = 0
        if end_time <= start_time:
            end_time = start_time + 5.0
        if video_duration:
            if start_time >= video_duration:
                start_time = max(0, video_duration - 5.0)
                end_time = video_duration
            elif end_time > video_duration:
                end_time = video_duration
        return float(start_time), float(end_time)

    def extract_clip_with_padding(self, video_path: str, timestamp: float, duration: float = None):
        """Returns the (start, end) the reference would hand to ffmpeg; no transcode is performed here."""
        return self.clip_interval(timestamp, duration)
