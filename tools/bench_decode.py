"""Host side of the frame feed, timed on the CPU: FrameExtractor.extract_frames (every sampled frame, seek + read +
BGR2RGB per frame, what /root/reference/src/services/frame_extractor.py:76-104 does) against
FrameExtractor.extract_window_middles (only the frame each sliding window embeds, left in BGR order for K1).
usage: python tools/bench_decode.py [frames=480] [height=1080] [width=1920]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from b200clip.services.frame_extractor import FrameExtractor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 480
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
path = os.path.join(tempfile.mkdtemp(prefix="b200clip_decode_"), "v.mp4")
vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (w, h))
rng = np.random.default_rng(0)
base = rng.integers(0, 256, (h // 8, w // 8, 3), dtype=np.uint8)
for i in range(n):
    f = cv2.resize(np.roll(base, i, axis=1), (w, h), interpolation=cv2.INTER_LINEAR)
    cv2.rectangle(f, ((7 * i) % (w - 200), h // 3), ((7 * i) % (w - 200) + 200, h // 3 + 150), (0, 0, 255), -1)
    vw.write(f)
vw.release()
fx = FrameExtractor()
t0 = time.perf_counter()
frames, stamps = fx.extract_frames(path)
t1 = time.perf_counter()
mid_idx, wts = fx.window_middles(len(frames), stamps)
got = fx.extract_window_middles(path, bgr=True)
t2 = time.perf_counter()
assert got is not None and got[1] == wts and np.array_equal(got[0][..., ::-1], frames[np.asarray(mid_idx)])
print(f"{n} frames of {w}x{h} (mp4v, {os.path.getsize(path) / 1e6:.1f} MB), {len(wts)} windows of {fx.window_size} / stride {fx.window_stride}, "
      f"{os.cpu_count()} host cores")
print(f"extract_frames          (all {len(frames)} sampled frames, RGB): {t1 - t0:7.3f} s, {frames.nbytes / 1e6:8.1f} MB of host frames")
print(f"extract_window_middles  ({len(wts)} middle frames + sentinel, BGR): {t2 - t1:7.3f} s, {got[0].nbytes / 1e6:8.1f} MB of host frames "
      f"({(t1 - t0) / (t2 - t1):.1f}x less decode time, same frames and timestamps)")
