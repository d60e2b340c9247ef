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
    from synth import write_test_video

    return write_test_video(path, n, w, h, fps)


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


def test_decode_semantics_equal_the_reference_frame_extractor(tmp_path, monkeypatch, golden_dir):
    """tests/golden/frame_extractor.json = the reference's own FrameExtractor (OpenCV path) on mp4 files written by
    synth.write_test_video (tests/golden/make_golden_frames.py): sampled-frame count (incl. the 1000-frame cap with
    step 1 and 2), timestamps from the decoder position, window timestamps, and -- after the reference's <= 512 x 512
    INTER_AREA shrink, which this repository performs inside K1 instead -- the very pixels of every kept frame."""
    import json
    import os
    import zlib

    from oracle.reference_pipeline import resize_frame_for_memory
    from synth import write_test_video

    cases = json.load(open(os.path.join(golden_dir, "frame_extractor.json")))["cases"]
    assert len(cases) >= 7
    for c in cases:
        fx = _extractor(monkeypatch, FRAME_SAMPLE_RATE=c["sample_rate"])
        path = write_test_video(tmp_path / f"v_{c['n']}_{c['w']}_{c['sample_rate']}.mp4", c["n"], c["w"], c["h"], c["fps"])
        frames, stamps = fx.extract_frames(path)
        assert stamps == c["timestamps"] and len(frames) == c["shape"][0]
        shrunk = [resize_frame_for_memory(f) for f in frames]
        assert list(shrunk[0].shape) == c["shape"][1:] and str(shrunk[0].dtype) == c["dtype"]
        assert [zlib.crc32(np.ascontiguousarray(f).tobytes()) for f in shrunk] == c["frame_crc"]
        mid_idx, wts = fx.window_middles(len(frames), stamps)
        assert wts == c["window_timestamps"] and len(wts) == c["windows"]
        middle, ts, n_sampled = fx.extract_window_middles(path)
        assert ts == c["window_timestamps"] and n_sampled == c["shape"][0]
        assert [zlib.crc32(np.ascontiguousarray(resize_frame_for_memory(f)).tobytes()) for f in middle] == [c["frame_crc"][i] for i in mid_idx]
