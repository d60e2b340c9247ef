"""Pins oracle/phase1_ref.py against outputs of the reference's own Phase1MVP / VideoProcessor / FrameExtractor
(tests/golden/phase1_cfg1.json, queries.json) and checks the host-side mirror (b200clip.services / pipeline logic
that needs no GPU) against the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import phase1_ref as R


def test_windows_and_topk_reproduce_reference_run(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "phase1_cfg1.json")))
    ts = [i / g["fps"] for i in range(g["n_frames"])]
    idx, wts = R.sliding_window_middles(g["n_frames"], ts)
    assert len(idx) == 63 and idx[0] == 8 and idx[-1] == 504 and wts == g["window_timestamps"]
    sims = np.array(g["similarities"], np.float32)
    assert R.topk_threshold(sims, wts, 5, -1.0) == g["top5"]
    assert R.topk_threshold(sims, wts, 63, -1.0) == g["all"]
    thr = g["threshold_case"]["threshold"]
    assert R.topk_threshold(sims, wts, 5, thr) == g["threshold_case"]["results"]
    assert R.topk_threshold(sims, wts, 5, 0.25) == []


def test_query_normalisation_matches_reference(golden_dir):
    q = json.load(open(os.path.join(golden_dir, "queries.json")))
    for raw, want in q.items():
        assert R.preprocess_query(raw) == want


@pytest.mark.parametrize("n", [0, 1, 5, 15, 16, 17, 23, 24, 512, 1000, 3600])
def test_host_mirror_windows_match_oracle(n):
    from b200clip.services.frame_extractor import FrameExtractor

    ts = [i / 30.0 for i in range(n)]
    fe = FrameExtractor()
    assert fe.window_middles(n, ts) == R.sliding_window_middles(n, ts)
    frames = np.arange(n, dtype=np.uint8).reshape(n, 1, 1, 1).repeat(3, 3)
    win, wts = fe.create_sliding_windows(frames, ts)
    idx, wts2 = R.sliding_window_middles(n, ts)
    assert list(wts) == wts2
    if n:
        assert [int(w[len(w) // 2][0, 0, 0]) for w in win] == [i % 256 for i in idx]
    with pytest.raises(ValueError):
        fe.window_middles(n + 1, ts)


def test_host_mirror_sampling_cap():
    from b200clip.services.frame_extractor import FrameExtractor

    fe = FrameExtractor()
    assert fe.sample_indices(500) == list(range(500))
    idx = fe.sample_indices(108000)            # 1 h at 30 fps -> capped at 1000, step 108
    assert len(idx) == 1000 and idx[1] - idx[0] == 108
    idx = fe.sample_indices(1999)
    assert len(idx) == 1000 and idx[1] == 1


@pytest.mark.parametrize("t,dur,vd", [(2.0, 30, None), (100.0, 30, 110.0), (100.0, 30, 90.0), (50.0, 0, None),
                                      (0.0, 30, 5.0), (14.99, 30, 1000.0), (300.0, 30, 100.0)])
def test_clip_interval_mirror(t, dur, vd):
    from b200clip.services.clip_extractor import ClipExtractor

    assert ClipExtractor().clip_interval(t, dur, vd) == R.clip_interval(t, dur, vd)
    s, e = R.clip_interval(t, dur, vd)
    assert 0 <= s and (vd is None or e <= max(vd, s + 5.0))


def test_query_mirror_and_settings_defaults(golden_dir):
    from b200clip.services.video_processor import VideoProcessor
    from b200clip.utils.config import Settings

    q = json.load(open(os.path.join(golden_dir, "queries.json")))
    vp = VideoProcessor.__new__(VideoProcessor)
    for raw, want in q.items():
        assert vp.preprocess_query(raw) == want
    s = Settings()
    assert (s.WINDOW_SIZE, s.WINDOW_STRIDE, s.BATCH_SIZE, s.TOP_K_RESULTS, s.CONFIDENCE_THRESHOLD, s.CLIP_DURATION,
            s.MAX_FRAME_WIDTH, s.OPENCLIP_MODEL, s.OPENCLIP_PRETRAINED) == (16, 8, 32, 15, 0.25, 30, 512, "ViT-B-32",
                                                                            "openai")


def test_merge_oracle_equals_global_sort():
    rng = np.random.default_rng(0)
    s = rng.standard_normal(1000).astype(np.float32)
    s[100:104] = s[7]
    k, g = 10, 4
    cs, ci = [], []
    for r in range(g):
        lo, hi = r * 250, (r + 1) * 250
        o = np.lexsort((np.arange(lo, hi), s[lo:hi]))[::-1][:k]
        cs.append(s[lo:hi][o])
        ci.append(o + lo)
    ms, mi = R.merge_topk_lists(np.stack(cs), np.stack(ci), k)
    want = np.lexsort((np.arange(1000), s))[::-1][:k]
    assert np.array_equal(mi, want) and np.array_equal(ms, s[want])


def test_clip_interval_equals_the_reference_clip_extractor(golden_dir):
    """tests/golden/clip_intervals.json = the reference's own ClipExtractor.extract_clip_with_padding -> extract_clip with
    ffmpeg replaced by a recorder (tests/golden/make_golden_clips.py): 52 (timestamp, clip duration, video duration)
    cases incl. the negative-start, empty-interval and past-the-end clamps.  The oracle restatement and the product's
    host mirror (the K4 kernel's float64 intervals are compared with the oracle in tests/test_gpu_topk.py) must agree."""
    import json
    import os

    from b200clip.services.clip_extractor import ClipExtractor
    from b200clip.utils.config import settings

    cases = json.load(open(os.path.join(golden_dir, "clip_intervals.json")))["cases"]
    assert len(cases) >= 50
    ce = ClipExtractor()
    for c in cases:
        dur = settings.CLIP_DURATION if c["duration"] is None else c["duration"]
        for fn in (R.clip_interval, ce.clip_interval):
            s, e = fn(c["timestamp"], dur, c["video_duration"])
            assert s == c["start"] and abs(e - c["end"]) <= 1e-9 and abs((e - s) - c["t"]) <= 1e-9, (c, s, e)
