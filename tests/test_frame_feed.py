"""Frame feed, host side: FrameExtractor.extract_window_middles decodes only the frame each sliding
window embeds.  Pinned against the reference-shaped path -- extract_frames (every sampled frame, seek + read per index
like /root/reference/src/services/frame_extractor.py:76-104) followed by the window arithmetic (:237-273) -- on mp4
files written here with OpenCV: same frames (bytes), same window timestamps, for the default 16 / 8 windows, other
sample rates, the 1000-frame cap, videos shorter than a window, and a file whose tail does not decode (where the
shortcut must refuse, because the reference's frame list -- and with it every window -- is shorter there)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _write_video(path, n, w=96, h=64, fps=8.0):
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    assert vw.isOpened()
    xs = np.arange(w)[None, :]
    for i in range(n):
        f = np.zeros((h, w, 3), np.uint8)
        f[:, :, 0] = (3 * i) % 256
        f[:, :, 1] = ((xs + 5 * i) % 256).astype(np.uint8)
        f[h // 4:h // 2, (2 * i) % (w - 16):(2 * i) % (w - 16) + 16, 2] = 255
        vw.write(f)
    vw.release()
    return str(path)


def _extractor(monkeypatch, **over):
    from b200clip.services.frame_extractor import FrameExtractor
    from b200clip.utils.config import settings

    for k, v in over.items():
        monkeypatch.setattr(settings, k, v)
    return FrameExtractor()


@pytest.mark.parametrize("n,over", [(80, {}), (83, {"FRAME_SAMPLE_RATE": 3}), (100, {"MAX_SAMPLED_FRAMES": 20}),
                                    (10, {}), (16, {}), (17, {"WINDOW_STRIDE": 1}), (1, {})])
def test_middles_only_decode_equals_full_decode(tmp_path, monkeypatch, n, over):
    fx = _extractor(monkeypatch, **over)
    path = _write_video(tmp_path / "v.mp4", n)
    frames, stamps = fx.extract_frames(path)
    mid_idx, window_ts = fx.window_middles(len(frames), stamps)
    got = fx.extract_window_middles(path)
    assert got is not None
    middle, ts, n_sampled = got
    assert n_sampled == len(frames) and ts == window_ts and len(middle) == len(mid_idx) >= 1
    assert middle.dtype == np.uint8 and np.array_equal(middle, frames[np.asarray(mid_idx)])
    # decoder order (BGR) on request: the same frames with the channel swap of frame_extractor.py:191 left to K1
    raw, ts_b, n_b = fx.extract_window_middles(path, bgr=True)
    assert ts_b == ts and n_b == n_sampled and np.array_equal(raw[..., ::-1], middle)
    # and the same windows as the reference-shaped window builder
    windows, wts = fx.create_sliding_windows(frames, stamps)
    assert wts == window_ts
    assert np.array_equal(np.stack([w[len(w) // 2] for w in windows]), middle)


def test_truncated_tail_refuses_the_shortcut(tmp_path, monkeypatch):
    """Frame count says 80, the decoder delivers 70: extract_frames returns 70 frames (windows over 70), so the
    shortcut -- which planned its windows over 80 -- must return None and leave the decision to the full path."""
    fx = _extractor(monkeypatch)
    path = _write_video(tmp_path / "v.mp4", 80)
    real = cv2.VideoCapture

    class Truncated:
        def __init__(self, p):
            self.c, self.pos = real(p), 0

        def isOpened(self):
            return self.c.isOpened()

        def get(self, prop):
            return self.c.get(prop)

        def set(self, prop, v):
            if prop == cv2.CAP_PROP_POS_FRAMES:
                self.pos = int(v)
            return self.c.set(prop, v)

        def read(self):
            if self.pos >= 70:
                return False, None
            self.pos += 1
            return self.c.read()

        def release(self):
            self.c.release()

    monkeypatch.setattr(cv2, "VideoCapture", Truncated)
    frames, stamps = fx.extract_frames(path)
    assert len(frames) == 70
    assert fx.extract_window_middles(path) is None


def test_unreadable_file_raises_like_extract_frames(tmp_path, monkeypatch):
    fx = _extractor(monkeypatch)
    bad = tmp_path / "x.mp4"
    bad.write_bytes(b"0")
    with pytest.raises(ValueError, match="Cannot open video"):
        fx.extract_window_middles(str(bad))
    with pytest.raises(ValueError, match="Cannot open video"):
        fx.extract_frames(str(bad))
