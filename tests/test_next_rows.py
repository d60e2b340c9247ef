"""SURVEY.md section 8f "next" rows on the CPU: the oracle restatements and the product's host code for temporal
consistency (phase3_advanced.py:37-81) and single-stage image matching (image_matcher.py:980-1018) against
tests/golden/next_rows.json = outputs of the reference's own functions (tests/golden/make_golden_next.py)."""
import json
import os

import numpy as np
import pytest

from oracle import phase1_ref as R


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "next_rows.json")))


def test_temporal_consistency_matches_reference(golden):
    from b200clip.pipeline.temporal import apply_temporal_consistency, merge_hits

    assert len(golden["temporal_consistency"]) == 40
    dropped = 0
    for case in golden["temporal_consistency"]:
        want = case["output"]
        assert R.temporal_consistency([dict(h) for h in case["input"]]) == want          # oracle == reference
        got = apply_temporal_consistency([dict(h) for h in case["input"]])
        assert got == want                                                                # product == reference
        dropped += len(case["input"]) - len(want)
        merged = merge_hits([dict(h) for h in case["input"]])
        assert sorted(map(json.dumps, merged)) == sorted(map(json.dumps, want))
        assert [m["confidence"] for m in merged] == sorted((m["confidence"] for m in merged), reverse=True)
    assert dropped > 20        # the fixture does exercise the suppression


def test_temporal_consistency_properties():
    from b200clip.pipeline.temporal import apply_temporal_consistency

    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(0, 12))
        hits = [{"timestamp": float(rng.uniform(0, 30)), "confidence": float(rng.uniform(0, 1))} for _ in range(n)]
        out = apply_temporal_consistency([dict(h) for h in hits])
        assert out == R.temporal_consistency([dict(h) for h in hits])
        assert all(o in hits for o in out) and len(out) <= len(hits)
        if hits:
            best = max(hits, key=lambda h: h["confidence"])
            # a hit is only ever dropped in favour of a strictly better or equal one, so the best survives unless tied
            assert best in out or sum(h["confidence"] == best["confidence"] for h in hits) > 1


def test_single_stage_ranking_matches_reference(golden):
    from b200clip.services.image_matcher import ImageMatcher

    assert len(golden["single_stage_matching"]) == 12
    ties = 0
    for case in golden["single_stage_matching"]:
        sims = np.array(case["similarities"], np.float32)
        want = case["output"]
        ref = R.single_stage_matching(case["similarities"], case["timestamps"], case["top_k"], case["threshold"])
        assert ref == want                                                                # oracle == reference
        got = ImageMatcher.rank_single_stage(sims, case["timestamps"], case["top_k"], case["threshold"])
        assert [g["frame_index"] for g in got] == [w["frame_index"] for w in want]
        assert [g["timestamp"] for g in got] == [w["timestamp"] for w in want]
        assert np.allclose([g["confidence"] for g in got], [w["confidence"] for w in want], atol=1e-6)
        assert all(set(g) == {"timestamp", "confidence", "clip_similarity", "method", "frame_index"} for g in got)
        conf = [w["confidence"] for w in want]
        ties += len(conf) - len(set(conf))
    assert ties > 0            # ties (lower frame index first) are covered


def _tiny_frames(n):
    f = np.zeros((n, 2, 2, 3), np.uint8)
    f[:, 0, 0, 0] = np.arange(n) % 256
    f[:, 0, 0, 1] = np.arange(n) // 256
    return f


def test_phase2_handoff_equals_the_reference_reranker(golden_dir):
    """tests/golden/phase2_handoff.json = the reference's own Phase2Reranker.process_video
    (/root/reference/src/pipeline/phase2_reranker.py:31-90) with phase 1, the frame extractor and BLIP replaced by
    deterministic stubs (tests/golden/make_golden_phase2.py).  The same stubs around OUR hand-off must give the same
    dicts: candidates asked of phase 1 (2 * top_k, 20 for None), captioned frame (window middle), 0.7 / 0.3 blend,
    stable descending sort, truncation, empty result, debug tuple."""
    import json

    from b200clip.pipeline.phase2_handoff import Phase2Reranker
    from b200clip.services.frame_extractor import FrameExtractor

    cases = json.load(open(os.path.join(golden_dir, "phase2_handoff.json")))["cases"]
    assert len(cases) >= 16
    for c in cases:
        n = c["n_frames"]
        frames, stamps = _tiny_frames(n), [round(i / c["fps"], 4) for i in range(n)]
        real = FrameExtractor()
        m = len(real.window_middles(n, stamps)[0])
        asked = []

        class P1:
            def process_video(self, video_path, query, top_k=None, debug_mode=None):
                asked.append(top_k)
                res = [dict(h) for h in c["hits"][:top_k]]
                return (res, [{"window_index": i} for i in range(m)]) if debug_mode else res

        class FX:
            def extract_frames(self, video_path):
                return frames, stamps

            def create_sliding_windows(self, fr, ts):
                return real.create_sliding_windows(fr, ts)

        class Blip:
            def generate_caption(self, frame):
                return f"frame {int(frame[0, 0, 0]) + 256 * int(frame[0, 0, 1])}"

            def compute_text_similarity(self, caption, query):
                return float(c["caption_table"][int(caption.split()[1])])

        r = Phase2Reranker(phase1=P1(), caption_model=Blip(), frame_extractor=FX())
        out = r.process_video("video.mp4", "red car", c["top_k"], debug_mode=c["debug"])
        assert asked == [c["asked_phase1_for"]]
        assert out == c["output"]
    with pytest.raises(RuntimeError, match="captioning model"):
        Phase2Reranker(phase1=object(), caption_model=None, frame_extractor=object()).process_video("v.mp4", "q", 3)


def test_phase2_candidate_frames_decode_only_what_is_captioned(tmp_path):
    """On a real mp4 the re-ranker's frame access (extract_window_middles(only=...)) must hand BLIP the very frames the
    reference's decode-everything + create_sliding_windows path would."""
    cv2 = pytest.importorskip("cv2")
    from b200clip.pipeline.phase2_handoff import Phase2Reranker
    from b200clip.services.frame_extractor import FrameExtractor

    path = str(tmp_path / "v.mp4")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 8.0, (96, 64))
    for i in range(90):
        f = np.zeros((64, 96, 3), np.uint8)
        f[:, :, 0] = (3 * i) % 256
        f[16:32, (2 * i) % 80:(2 * i) % 80 + 16, 2] = 255
        vw.write(f)
    vw.release()
    fx = FrameExtractor()
    frames, stamps = fx.extract_frames(path)
    windows, wts = fx.create_sliding_windows(frames, stamps)
    want = [7, 2, 9, 2, 0]
    r = Phase2Reranker(phase1=object(), caption_model=object(), frame_extractor=fx)
    got = r._middle_frames(path, want)
    assert sorted(got) == [0, 2, 7, 9]
    for w in got:
        assert np.array_equal(got[w], windows[w][len(windows[w]) // 2])
    sel, ts, n_sampled = fx.extract_window_middles(path, only=[9, 0])
    assert ts == [wts[9], wts[0]] and n_sampled == len(frames) and np.array_equal(sel[0], windows[9][8])
    with pytest.raises(IndexError):
        fx.extract_window_middles(path, only=[len(wts)])


def test_process_query_reranked_mode_uses_the_injected_phase2(tmp_path):
    """video_processor.py:432-447: "reranked" / "advanced" go through phase 2 when it is available and fall back to
    phase 1 otherwise; the threshold filter and the clip intervals apply to whichever list comes back."""
    from b200clip.services.video_processor import VideoProcessor

    f = tmp_path / "v.mp4"
    f.write_bytes(b"0")
    calls = []

    class P1:
        def process_video(self, video_path, query, top_k=None, debug_mode=None, merge=None):
            calls.append(("p1", query, top_k))
            return [{"timestamp": 4.0, "confidence": 0.5, "phase": "phase1_mvp", "window_index": 1}]

    class P2:
        def process_video(self, video_path, query, top_k=None, debug_mode=False):
            calls.append(("p2", query, top_k))
            return [{"timestamp": 40.0, "confidence": 0.6, "phase": "phase2_reranked", "window_index": 3, "caption": "c",
                     "clip_score": 0.5, "caption_score": 0.83},
                    {"timestamp": 2.0, "confidence": 0.1, "phase": "phase2_reranked", "window_index": 0, "caption": "d",
                     "clip_score": 0.1, "caption_score": 0.1}]

    vp = VideoProcessor(phase1=P1(), phase2=P2())
    out = vp.process_query(str(f), "A dog  jumps", mode="reranked", top_k=4, threshold=0.25)
    assert calls == [("p2", "dog jumping", 4)] and out["status"] == "success" and out["mode"] == "reranked"
    assert [r["phase"] for r in out["results"]] == ["phase2_reranked"] and out["total_found"] == 1
    assert out["results"][0]["clip_start"] == 25.0 and out["results"][0]["clip_end"] == 55.0
    assert vp.process_query(str(f), "dog", mode="advanced", threshold=0.0)["total_found"] == 2
    calls.clear()
    fallback = VideoProcessor(phase1=P1()).process_query(str(f), "dog", mode="reranked", top_k=2, threshold=0.25)
    assert calls == [("p1", "dog", 2)] and fallback["results"][0]["phase"] == "phase1_mvp"
    assert vp.process_query(str(f), "dog", mode="mvp", top_k=2)["results"][0]["phase"] == "phase1_mvp"


def test_image_matcher_stage2_clip_filter(monkeypatch):
    """image_matcher.py:407-415 on a table of similarities: every candidate gets its clip_similarity, the list is cut at
    thresholds['clip_similarity'] (>=), order and the other keys are kept; frames of different shapes are batched per shape."""
    from b200clip.services.image_matcher import ImageMatcher

    # (5: float32(0.7) = 0.69999999 is below the Python-float threshold 0.7 -- also in the reference, whose
    # similarities are float32 cosines converted with float())
    table = {3: 0.71, 5: 0.7, 9: 0.7001, 12: 0.95, 40: 0.2}
    m = ImageMatcher(clip_model=object())
    batches = []

    def fake(reference_image, frames):
        batches.append(frames.shape)
        return np.array([table[int(f[0, 0, 0])] for f in frames], np.float32)

    monkeypatch.setattr(m, "clip_similarities", fake)
    cands = []
    for i, shape in ((3, (4, 4, 3)), (5, (4, 4, 3)), (9, (2, 6, 3)), (12, (4, 4, 3)), (40, (2, 6, 3))):
        f = np.zeros(shape, np.uint8)
        f[0, 0, 0] = i
        cands.append({"index": i, "timestamp": i / 2.0, "frame": f, "hash_distance": 7})
    kept = m.stage2_clip_filter(np.zeros((8, 8, 3), np.uint8), cands)
    assert [c["index"] for c in kept] == [3, 9, 12]
    assert [c["clip_similarity"] for c in kept] == [float(np.float32(0.71)), float(np.float32(0.7001)), float(np.float32(0.95))]
    assert cands[1]["clip_similarity"] == float(np.float32(0.7)) and cands[4]["clip_similarity"] == float(np.float32(0.2))
    assert all(c["hash_distance"] == 7 for c in kept) and sorted(batches) == [(2, 2, 6, 3), (3, 4, 4, 3)]
    assert m.stage2_clip_filter(np.zeros((8, 8, 3), np.uint8), []) == []
