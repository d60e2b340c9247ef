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
