"""CPU tests of the .b2emb embedding-cache format (services/embedding_cache.py): round trips, bf16 rounding equal
to torch's, block commits, resume after a torn write, header validation."""
import os
import struct

import numpy as np
import pytest
import torch

from b200clip.services import embedding_cache as EC


def _rows(n, e, seed):
    x = np.random.default_rng(seed).standard_normal((n, e)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_round_trip_and_blocks(tmp_path, dtype):
    p = str(tmp_path / "v.b2emb")
    a, b = _rows(70, 64, 1), _rows(33, 64, 2)
    with EC.EmbeddingCacheWriter(p, 64, dtype, True, {"model": "ViT-B-32", "weights_fingerprint": "abc"}, 12.5, 2.0) as w:
        w.append(a, np.arange(70) / 2.0)
        w.append(torch.from_numpy(b), 35.0 + np.arange(33) / 2.0)
        w.append(np.empty((0, 64), np.float32), [])
        assert w.rows == 103
    h, ts, rows = EC.read_cache(p)
    assert (h.embed_dim, h.rows, h.duration, h.fps, h.window_size, h.window_stride) == (64, 103, 12.5, 2.0, 16, 8)
    assert h.meta["model"] == "ViT-B-32" and h.flags & EC.FLAG_UNIT_NORM
    assert np.array_equal(ts, np.concatenate([np.arange(70) / 2.0, 35.0 + np.arange(33) / 2.0]))
    want = np.concatenate([a, b])
    if dtype == "float32":
        assert rows.dtype == np.float32 and np.array_equal(rows, want)
    else:
        assert rows.dtype == np.uint16
        tb = torch.from_numpy(want).bfloat16().view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(rows, tb)                      # same rounding as torch (nearest even)
        assert np.abs(EC._f32_from_bf16_bits(rows) - want).max() < 4e-3
    assert os.path.getsize(p) % 64 == 0


def test_bf16_tensor_is_written_verbatim(tmp_path):
    p = str(tmp_path / "b.b2emb")
    t = torch.from_numpy(_rows(9, 32, 3)).bfloat16()
    with EC.EmbeddingCacheWriter(p, 32, "bfloat16") as w:
        w.append(t, range(9))
    _, _, rows = EC.read_cache(p)
    assert np.array_equal(rows, t.view(torch.int16).numpy().view(np.uint16))


def test_resume_after_torn_block(tmp_path):
    p = str(tmp_path / "r.b2emb")
    a, b, c = _rows(40, 16, 4), _rows(25, 16, 5), _rows(10, 16, 6)
    with EC.EmbeddingCacheWriter(p, 16, "float32", meta={"weights_fingerprint": "w1"}) as w:
        w.append(a, range(40))
        w.append(b, range(40, 65))
    good = os.path.getsize(p)
    with open(p, "ab") as f:                                  # a crash in the middle of the next block
        f.write(struct.pack("<Q", 10) + b"\x01" * 333)
    h, ts, rows = EC.read_cache(p)                            # readers only ever see committed rows
    assert h.rows == 65 and np.array_equal(rows, np.concatenate([a, b]))
    with EC.EmbeddingCacheWriter(p, 16, "float32", meta={"weights_fingerprint": "w1"}, resume=True) as w:
        assert w.rows == 65 and os.path.getsize(p) == good    # the torn tail is gone
        w.append(c, range(65, 75))
    h, ts, rows = EC.read_cache(p)
    assert h.rows == 75 and np.array_equal(rows, np.concatenate([a, b, c])) and list(ts) == list(range(75))
    with pytest.raises(ValueError):
        EC.EmbeddingCacheWriter(p, 16, "float32", meta={"weights_fingerprint": "other"}, resume=True)
    with pytest.raises(ValueError):
        EC.EmbeddingCacheWriter(p, 32, "float32", resume=True)


def test_header_validation(tmp_path):
    p = str(tmp_path / "x.b2emb")
    open(p, "wb").write(b"not a cache file" * 8)
    with pytest.raises(ValueError):
        EC.read_cache(p)
    with EC.EmbeddingCacheWriter(p, 8, "float32") as w:
        w.append(_rows(3, 8, 7), [0, 1, 2])
    raw = bytearray(open(p, "rb").read())
    raw[24:32] = struct.pack("<Q", 999)                       # claims rows that are not there
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        EC.read_cache(p)
    with pytest.raises(ValueError):
        with EC.EmbeddingCacheWriter(str(tmp_path / "y.b2emb"), 8) as w:
            w.append(_rows(3, 9, 8), [0, 1, 2])


def test_cache_path_is_keyed_by_file_model_and_weights(tmp_path):
    v = tmp_path / "clip.mp4"
    v.write_bytes(b"0" * 100)
    a = EC.cache_path_for(str(v), str(tmp_path), "ViT-B-32", "w1")
    assert a == EC.cache_path_for(str(v), str(tmp_path), "ViT-B-32", "w1") and a.endswith(".b2emb")
    assert a != EC.cache_path_for(str(v), str(tmp_path), "ViT-L-14", "w1")
    assert a != EC.cache_path_for(str(v), str(tmp_path), "ViT-B-32", "w2")
    sd = {"a": torch.ones(3, 4), "b": torch.zeros(5)}
    assert EC.weights_fingerprint(sd) == EC.weights_fingerprint(dict(reversed(list(sd.items()))))
    assert EC.weights_fingerprint(sd) != EC.weights_fingerprint({"a": torch.ones(3, 4) * 2, "b": torch.zeros(5)})
