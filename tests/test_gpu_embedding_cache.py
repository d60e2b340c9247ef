"""Embedding cache + multi-query service on the GPU: a cached video answers queries
exactly like Phase1MVP.process_frames on the same frames, batched queries equal single ones, files round-trip and
an interrupted build resumes."""
import numpy as np
import pytest
import torch

from synth import structured_frames

pytestmark = pytest.mark.gpu

QUERIES = ["a person walking across street", "red car", "dog running on grass", "two people talking"]


@pytest.fixture(scope="module")
def phase1(oracle_sd_b32):
    from b200clip.models.openclip_model import OpenCLIPModel
    from b200clip.pipeline.phase1_mvp import Phase1MVP
    from b200clip.utils.config import settings

    settings.B200_MAX_IMAGES_PER_PASS = 128
    return Phase1MVP(clip_model=OpenCLIPModel(state_dict=oracle_sd_b32))


def test_cached_queries_equal_phase1(phase1, tmp_path):
    from b200clip.services.embedding_cache import EmbeddingCache, read_cache
    from b200clip.utils.config import settings

    video = structured_frames(200, 224, 224, seed=5)
    ts = [i / 4.0 for i in range(len(video))]
    settings.CONFIDENCE_THRESHOLD = -1.0
    try:
        cache = EmbeddingCache.build(phase1.clip_model, video, ts, dtype="float32", duration=50.0,
                                     path=str(tmp_path / "v.b2emb"), chunk=7)
        assert len(cache) == 24 and read_cache(str(tmp_path / "v.b2emb"))[0].rows == 24
        batch = cache.query_batch(QUERIES, top_k=5)
        for q, got in zip(QUERIES, batch):
            want = phase1.process_frames(video, ts, q, top_k=5, video_duration=50.0)
            assert [r["window_index"] for r in got] == [r["window_index"] for r in want]
            assert [r["confidence"] for r in got] == [r["confidence"] for r in want]
            assert [r["timestamp"] for r in got] == [r["timestamp"] for r in want]
            assert got == cache.query(q, top_k=5)
            for r in got:       # clip_extractor.py:175-183 on a 50 s video
                assert r["start"] == max(0.0, r["timestamp"] - 15.0) and r["end"] == min(50.0, r["timestamp"] + 15.0)
        # file round trip, and the bf16 form of the same cache
        again = EmbeddingCache.load(phase1.clip_model, str(tmp_path / "v.b2emb"))
        assert torch.equal(again.embeddings, cache.embeddings) and again.timestamps == cache.timestamps
        assert again.query_batch(QUERIES, top_k=5) == batch
        half = EmbeddingCache(phase1.clip_model, cache.embeddings.bfloat16(), cache.timestamps, 50.0)
        half.save(str(tmp_path / "h.b2emb"))
        half2 = EmbeddingCache.load(phase1.clip_model, str(tmp_path / "h.b2emb"))
        assert torch.equal(half2.embeddings, half.embeddings)
        for a, b in zip(half2.query_batch(QUERIES, top_k=5), batch):
            assert np.abs(np.array([r["confidence"] for r in a]) - np.array([r["confidence"] for r in b])).max() < 1e-2
        # thresholding happens on the device: a high threshold empties the lists
        assert cache.query_batch(QUERIES, top_k=5, threshold=0.99) == [[], [], [], []]
    finally:
        settings.CONFIDENCE_THRESHOLD = 0.25


def test_interrupted_build_resumes(phase1, tmp_path):
    from b200clip.services.embedding_cache import EmbeddingCache, EmbeddingCacheWriter, read_cache

    video = structured_frames(120, 224, 224, seed=6)
    ts = [i / 2.0 for i in range(len(video))]
    full = EmbeddingCache.build(phase1.clip_model, video, ts, dtype="bfloat16", chunk=5)
    p = str(tmp_path / "part.b2emb")
    with EmbeddingCacheWriter(p, full.embeddings.shape[1], "bfloat16") as w:      # an earlier run got 10 windows far
        w.append(full.embeddings[:10], full.timestamps[:10])
    launches_before = phase1.clip_model.model.handle.launches
    resumed = EmbeddingCache.build(phase1.clip_model, video, ts, dtype="bfloat16", chunk=5, path=p, resume=True)
    assert torch.equal(resumed.embeddings, full.embeddings) and resumed.timestamps == full.timestamps
    assert read_cache(p)[0].rows == len(full)
    # nothing to do the third time: no kernel is launched
    launches_before = phase1.clip_model.model.handle.launches
    EmbeddingCache.build(phase1.clip_model, video, ts, dtype="bfloat16", chunk=5, path=p, resume=True)
    assert phase1.clip_model.model.handle.launches == launches_before


def test_many_queries_over_a_large_cache_use_the_tensor_core_path(phase1):
    """>= 8 queries over a bf16 cache of >= 4096 rows: one tcgen05 K4 pass; planted windows come back per query."""
    from b200clip.services.embedding_cache import EmbeddingCache

    n, e = 6000, phase1.clip_model.model.embed_dim
    g = torch.Generator(device="cuda").manual_seed(1)
    emb = torch.randn(n, e, device="cuda", generator=g)
    emb = emb / emb.norm(dim=-1, keepdim=True)
    queries = [f"query number {i}" for i in range(12)]
    txt = torch.from_numpy(phase1.clip_model.encode_text(queries)).cuda()
    for qi in range(12):
        emb[100 + 37 * qi] = txt[qi]                       # cosine 1 with its own query
    cache = EmbeddingCache(phase1.clip_model, emb.bfloat16(), [i / 10.0 for i in range(n)], duration=600.0)
    res = cache.query_batch(queries, top_k=3, threshold=0.5)
    planted = {100 + 37 * qi for qi in range(12)}
    for qi, r in enumerate(res):
        # random-init text embeddings of different strings are strongly correlated, so other queries' planted rows
        # may follow; the query's own row (cosine 1) must lead and nothing but planted rows can pass 0.5
        assert 1 <= len(r) <= 3 and r[0]["window_index"] == 100 + 37 * qi and r[0]["confidence"] > 0.99
        assert r[0]["timestamp"] == (100 + 37 * qi) / 10.0
        assert all(x["window_index"] in planted and x["confidence"] >= 0.5 for x in r)
        assert [x["confidence"] for x in r] == sorted((x["confidence"] for x in r), reverse=True)


def test_process_video_opt_in_cache(phase1, tmp_path, monkeypatch):
    """settings.B200_EMBEDDING_CACHE: the first query embeds the video into DATA_DIR/embeddings, later queries never
    decode again; results are identical to the uncached path."""
    cv2 = pytest.importorskip("cv2")
    from pathlib import Path

    from b200clip.utils.config import settings

    video = structured_frames(64, 224, 224, seed=8)
    path = str(tmp_path / "clip.mp4")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 8.0, (224, 224))
    if not w.isOpened():
        pytest.skip("no mp4v encoder in this OpenCV build")
    for f in video:
        w.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    w.release()
    monkeypatch.setattr(settings, "CONFIDENCE_THRESHOLD", -1.0)
    monkeypatch.setattr(settings, "DATA_DIR", Path(tmp_path))
    want = phase1.process_video(path, "red car", top_k=5)
    assert len(want) == 5
    monkeypatch.setattr(settings, "B200_EMBEDDING_CACHE", True)
    got = phase1.process_video(path, "red car", top_k=5)
    assert got == want
    files = list((tmp_path / "embeddings").glob("clip-*.b2emb"))
    assert len(files) == 1
    # later queries (also from a fresh Phase1MVP sharing the model) are served from the file: no decode
    from b200clip.pipeline.phase1_mvp import Phase1MVP

    fresh = Phase1MVP(clip_model=phase1.clip_model)
    monkeypatch.setattr(fresh.frame_extractor, "extract_frames", lambda p: (_ for _ in ()).throw(AssertionError("decoded")))
    monkeypatch.setattr(fresh.frame_extractor, "extract_window_middles", lambda p: (_ for _ in ()).throw(AssertionError("decoded")))
    assert fresh.process_video(path, "red car", top_k=5) == want
    other = fresh.process_video(path, "dog running on grass", top_k=3)
    monkeypatch.setattr(settings, "B200_EMBEDDING_CACHE", False)
    assert other == phase1.process_video(path, "dog running on grass", top_k=3)
