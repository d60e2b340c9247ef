"""Embedding cache + multi-query service.

The reference reserves `data/embeddings/` for cached embeddings (/root/reference/README.md:208,
data/embeddings/.gitkeep) but never implements it: every query re-decodes and re-embeds the video
(src/pipeline/phase1_mvp.py:36-131).  Here a video is embedded ONCE (K1 -> K3), the window embeddings and their
timestamps are kept -- on the device and, optionally, in a `.b2emb` file -- and any number of queries is answered by
K4 alone (similarity + top-k + threshold + clip intervals; the tensor-core kernel from 8 queries on).  Results are
exactly what `Phase1MVP.process_video` returns for the same video and query.

File format `.b2emb` (little endian), designed so that a partially written file is still usable (resume):

    offset  size  field
    0       8     magic  b"B2EMB\\x01\\x00\\x00"
    8       4     header_bytes (u32)  -- offset of the first record block, multiple of 64
    12      4     embed_dim E (u32)
    16      4     dtype (u32): 0 = float32, 1 = bfloat16
    20      4     flags (u32): bit 0 = rows are unit L2 norm
    24      8     rows_committed N (u64)     -- rewritten after every appended block (the commit point)
    32      8     video_duration seconds (f64), 0 = unknown
    40      8     fps used for sampling (f64), 0 = unknown
    48      4     window_size (u32), 52: window_stride (u32)   -- frame_extractor.py:237-273 parameters
    56      4     meta_bytes (u32), 60: reserved
    64      ...   meta: UTF-8 JSON {model, pretrained, weights_fingerprint, source, ...}, zero padded to header_bytes
    then record blocks, each: u64 n_rows, f64 timestamps[n_rows], payload[n_rows * E] (dtype), zero padded to 64 bytes

`rows_committed` only counts whole blocks whose bytes are on disk, so a crash while appending loses at most the block
being written; `EmbeddingCacheWriter(path, resume=True)` truncates the file back to the last committed block and
continues from `rows_committed`."""
from __future__ import annotations

import hashlib
import json
import os
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

MAGIC = b"B2EMB\x01\x00\x00"
# settings that decide WHICH rows a cache holds: part of the cache key, of the header meta, and validated on load
SAMPLING_KEYS = ("frame_sample_rate", "max_sampled_frames", "resize_mode", "model", "pretrained")
DTYPE_F32, DTYPE_BF16 = 0, 1
FLAG_UNIT_NORM = 1
_FIXED = struct.Struct("<8sIIIIQddIIII")     # 64 bytes


def _pad64(n: int) -> int:
    return (n + 63) & ~63


def weights_fingerprint(state_dict) -> str:
    """Short, order-independent digest of a state dict (embeddings of different weights must not be mixed)."""
    h = hashlib.sha256()
    for k in sorted(state_dict):
        v = state_dict[k]
        a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        h.update(k.encode())
        h.update(str(a.shape).encode())
        h.update(np.ascontiguousarray(a).view(np.uint8)[:: max(1, a.nbytes // 4096)].tobytes())
    return h.hexdigest()[:16]


def _bf16_bits_from_f32(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit patterns (uint16), round to nearest even -- what torch's .bfloat16() does."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = 0x7FC0
    return r


def _f32_from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


class CacheHeader:
    def __init__(self, embed_dim: int, dtype: int, flags: int, rows: int, duration: float, fps: float, window_size: int,
                 window_stride: int, meta: dict, header_bytes: int):
        self.embed_dim, self.dtype, self.flags, self.rows = embed_dim, dtype, flags, rows
        self.duration, self.fps, self.window_size, self.window_stride = duration, fps, window_size, window_stride
        self.meta, self.header_bytes = meta, header_bytes

    @property
    def elem_bytes(self) -> int:
        return 2 if self.dtype == DTYPE_BF16 else 4

    def pack(self) -> bytes:
        mj = json.dumps(self.meta, sort_keys=True).encode()
        hb = self.header_bytes or _pad64(_FIXED.size + len(mj))
        if _FIXED.size + len(mj) > hb:
            raise ValueError("metadata does not fit the reserved header")
        self.header_bytes = hb
        fixed = _FIXED.pack(MAGIC, hb, self.embed_dim, self.dtype, self.flags, self.rows, self.duration, self.fps,
                            self.window_size, self.window_stride, len(mj), 0)
        return fixed + mj + b"\0" * (hb - _FIXED.size - len(mj))

    @classmethod
    def unpack(cls, buf: bytes) -> "CacheHeader":
        if len(buf) < _FIXED.size:
            raise ValueError("not a .b2emb file (too short)")
        magic, hb, e, dt, fl, rows, dur, fps, ws, wst, mb, _ = _FIXED.unpack(buf[:_FIXED.size])
        if magic != MAGIC:
            raise ValueError("not a .b2emb file (bad magic)")
        if dt not in (DTYPE_F32, DTYPE_BF16) or e == 0 or hb % 64 or hb < _FIXED.size + mb:
            raise ValueError("corrupt .b2emb header")
        meta = json.loads(buf[_FIXED.size:_FIXED.size + mb].decode() or "{}") if len(buf) >= _FIXED.size + mb else {}
        return cls(e, dt, fl, rows, dur, fps, ws, wst, meta, hb)


class EmbeddingCacheWriter:
    """Append-only writer; every `append` is one committed block (checkpoint).  Use as a context manager."""

    def __init__(self, path: str, embed_dim: int, dtype: str = "bfloat16", unit_norm: bool = True, meta: dict = None,
                 duration: float = 0.0, fps: float = 0.0, window_size: int = 16, window_stride: int = 8,
                 resume: bool = False):
        self.path = path
        want_dt = DTYPE_BF16 if dtype in ("bfloat16", "bf16") else DTYPE_F32
        if resume and os.path.exists(path):
            with open(path, "rb") as f:
                head = f.read(1 << 16)
            self.header = CacheHeader.unpack(head)
            if self.header.embed_dim != embed_dim or self.header.dtype != want_dt:
                raise ValueError("resume: embed_dim / dtype differ from the existing cache")
            if meta and self.header.meta.get("weights_fingerprint") not in (None, meta.get("weights_fingerprint")):
                raise ValueError("resume: the existing cache was written with different weights")
            if self.header.window_size != window_size or self.header.window_stride != window_stride:
                raise ValueError("resume: the existing cache was built with other window settings")
            for key in SAMPLING_KEYS:       # frames sampled / resized differently are different rows
                if meta and key in meta and key in self.header.meta and self.header.meta[key] != meta[key]:
                    raise ValueError(f"resume: the existing cache was built with another {key}")
            self._f = open(path, "r+b")
            end = _scan_blocks(self._f, self.header)[1]
            self._f.truncate(end)            # drop a torn block behind the last commit
            self._f.seek(end)
        else:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            self.header = CacheHeader(embed_dim, want_dt, FLAG_UNIT_NORM if unit_norm else 0, 0, duration, fps,
                                      window_size, window_stride, dict(meta or {}), 0)
            # reserve room so that metadata can be rewritten in place later (duration is often known last)
            self.header.header_bytes = _pad64(_FIXED.size + len(json.dumps(self.header.meta)) + 256)
            self._f = open(path, "w+b")
            self._f.write(self.header.pack())
            self._f.flush()

    @property
    def rows(self) -> int:
        return self.header.rows

    def append(self, embeddings, timestamps: Sequence[float]):
        """embeddings: [n, E] float32 numpy / torch tensor (any device; bf16 tensors are written as they are)."""
        emb = embeddings
        if hasattr(emb, "detach"):
            import torch

            emb = emb.detach()
            if emb.dtype == torch.bfloat16:
                bits = emb.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
                emb = bits if self.header.dtype == DTYPE_BF16 else _f32_from_bf16_bits(bits)
            else:
                emb = emb.float().cpu().numpy()
        emb = np.asarray(emb)
        ts = np.asarray(timestamps, dtype=np.float64)
        if emb.ndim != 2 or emb.shape[1] != self.header.embed_dim or len(ts) != len(emb):
            raise ValueError(f"append: expected [n,{self.header.embed_dim}] rows with n timestamps")
        if len(emb) == 0:
            return
        if self.header.dtype == DTYPE_BF16:
            payload = emb if emb.dtype == np.uint16 else _bf16_bits_from_f32(emb)
        else:
            payload = emb.astype(np.float32, copy=False)
        block = struct.pack("<Q", len(emb)) + ts.tobytes() + np.ascontiguousarray(payload).tobytes()
        block += b"\0" * (_pad64(len(block)) - len(block))
        self._f.write(block)
        self._f.flush()
        os.fsync(self._f.fileno())
        # commit point: the row count in the header moves only after the block is durable
        self.header.rows += len(emb)
        self._f.seek(24)
        self._f.write(struct.pack("<Q", self.header.rows))
        self._f.flush()
        os.fsync(self._f.fileno())
        self._f.seek(0, os.SEEK_END)

    def set_duration(self, duration: float):
        self.header.duration = float(duration)
        self._f.seek(32)
        self._f.write(struct.pack("<d", self.header.duration))
        self._f.flush()
        os.fsync(self._f.fileno())
        self._f.seek(0, os.SEEK_END)

    def close(self):
        if self._f:
            self._f.flush()
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _scan_blocks(f, header: CacheHeader) -> Tuple[List[Tuple[int, int]], int]:
    """[(offset, n_rows)] of the committed blocks and the file offset right behind the last of them."""
    blocks, off, left = [], header.header_bytes, header.rows
    f.seek(0, os.SEEK_END)
    size = f.tell()
    while left > 0:
        f.seek(off)
        raw = f.read(8)
        if len(raw) < 8:
            raise ValueError("corrupt .b2emb file: committed rows are missing")
        (n,) = struct.unpack("<Q", raw)
        nbytes = _pad64(8 + n * 8 + n * header.embed_dim * header.elem_bytes)
        if n == 0 or n > left or off + nbytes > size:
            raise ValueError("corrupt .b2emb file: block does not match the committed row count")
        blocks.append((off, n))
        off += nbytes
        left -= n
    return blocks, off


def read_cache(path: str) -> Tuple[CacheHeader, np.ndarray, np.ndarray]:
    """-> (header, timestamps float64 [N], rows: uint16 bf16 bit patterns or float32, [N, E]).  Only committed rows."""
    with open(path, "rb") as f:
        header = CacheHeader.unpack(f.read(1 << 16))
        blocks, _ = _scan_blocks(f, header)
        e, eb = header.embed_dim, header.elem_bytes
        ts = np.empty(header.rows, np.float64)
        rows = np.empty((header.rows, e), np.uint16 if header.dtype == DTYPE_BF16 else np.float32)
        r = 0
        for off, n in blocks:
            f.seek(off + 8)
            ts[r:r + n] = np.frombuffer(f.read(n * 8), np.float64)
            rows[r:r + n] = np.frombuffer(f.read(n * e * eb), rows.dtype).reshape(n, e)
            r += n
    return header, ts, rows


class EmbeddingCache:
    """Device-resident window embeddings of one video + the multi-query front end on top of K4."""

    def __init__(self, clip_model, embeddings, timestamps: Sequence[float], duration: float = 0.0, meta: dict = None):
        import torch

        self.clip_model = clip_model
        self.model = clip_model.model
        self.embeddings = embeddings.to(self.model.device).contiguous()
        if self.embeddings.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("embeddings must be float32 or bfloat16")
        self.timestamps = [float(t) for t in timestamps]
        self._ts_dev = torch.tensor(self.timestamps, dtype=torch.float64, device=self.model.device)
        self.duration = float(duration)
        self.meta = dict(meta or {})

    # ---------------------------------------------------------------------------------------- construction
    @classmethod
    def build(cls, clip_model, frames: np.ndarray, timestamps: Sequence[float], dtype: str = "bfloat16",
              duration: float = 0.0, path: Optional[str] = None, chunk: int = 4096, resume: bool = False,
              resize_mode: Optional[int] = None, fingerprint: Optional[str] = None) -> "EmbeddingCache":
        """Embed the sliding-window middle frames of a decoded video (frame_extractor.py:237-273 semantics) in
        chunks; with `path` every chunk is appended to a .b2emb file as it completes, and `resume=True` skips the
        windows an earlier, interrupted run already committed."""
        from .frame_extractor import FrameExtractor

        mid_idx, window_ts = FrameExtractor().window_middles(len(frames), list(timestamps))
        fr = np.asarray(frames)
        return cls._build(clip_model, lambda lo, hi: fr[np.asarray(mid_idx[lo:hi], dtype=np.int64)], window_ts, dtype, duration,
                          path, chunk, resume, resize_mode, fingerprint)

    @classmethod
    def build_from_middles(cls, clip_model, middle_frames: np.ndarray, window_timestamps: Sequence[float],
                           dtype: str = "bfloat16", duration: float = 0.0, path: Optional[str] = None, chunk: int = 4096,
                           resume: bool = False, resize_mode: Optional[int] = None,
                           fingerprint: Optional[str] = None) -> "EmbeddingCache":
        """`build` for a caller that decoded only the frame each window embeds
        (FrameExtractor.extract_window_middles): one row per given frame, same file, same rows."""
        if len(middle_frames) != len(window_timestamps):
            raise ValueError(f"Frames and timestamps length mismatch: {len(middle_frames)} vs {len(window_timestamps)}")
        fr = np.asarray(middle_frames)
        return cls._build(clip_model, lambda lo, hi: fr[lo:hi], list(window_timestamps), dtype, duration, path, chunk,
                          resume, resize_mode, fingerprint)

    @classmethod
    def _build(cls, clip_model, middles_of, window_ts, dtype, duration, path, chunk, resume, resize_mode, fingerprint):
        import torch

        from .. import capi
        from ..utils.config import settings

        model = clip_model.model
        mid_idx = range(len(window_ts))
        tdt = torch.bfloat16 if dtype in ("bfloat16", "bf16") else torch.float32
        mode = capi.RESIZE_REFERENCE if resize_mode is None else resize_mode
        meta = {"model": settings.OPENCLIP_MODEL, "pretrained": settings.OPENCLIP_PRETRAINED, "windows": len(mid_idx),
                "frame_sample_rate": settings.FRAME_SAMPLE_RATE, "max_sampled_frames": settings.MAX_SAMPLED_FRAMES,
                "resize_mode": int(mode) & 0xff}          # (without the input-format flag capi.INPUT_BGR)
        if fingerprint:
            meta["weights_fingerprint"] = fingerprint
        writer, done, parts = None, 0, []
        if path:
            writer = EmbeddingCacheWriter(path, model.embed_dim, dtype, True, meta, duration, 0.0,
                                          settings.WINDOW_SIZE, settings.WINDOW_STRIDE, resume=resume)
            done = writer.rows
            if done:
                _, _, rows = read_cache(path)
                prev = torch.from_numpy(rows.view(np.int16) if rows.dtype == np.uint16 else rows).to(model.device)
                parts.append(prev.view(torch.bfloat16) if rows.dtype == np.uint16 else prev)
        for lo in range(done, len(mid_idx), chunk):
            hi = min(lo + chunk, len(mid_idx))
            batch = np.ascontiguousarray(middles_of(lo, hi))
            if batch.dtype != np.uint8:
                batch = (batch * 255).astype(np.uint8)
            emb = torch.empty(hi - lo, model.embed_dim, device=model.device, dtype=torch.float32)
            model.encode_frames_u8_host(batch, resize_mode=mode, normalize=True, out=emb)
            emb = emb.to(tdt)
            parts.append(emb)
            if writer:
                writer.append(emb, window_ts[lo:hi])
        if writer:
            writer.close()
        all_emb = torch.cat(parts) if parts else torch.empty(0, model.embed_dim, device=model.device, dtype=tdt)
        return cls(clip_model, all_emb.to(tdt), window_ts, duration, meta)

    @classmethod
    def load(cls, clip_model, path: str, expect: Optional[dict] = None) -> "EmbeddingCache":
        """`expect`: {weights_fingerprint, window_size, window_stride, frame_sample_rate, ...} of the CURRENT settings;
        a cache written under different ones raises ValueError (the caller rebuilds)."""
        import torch

        header, ts, rows = read_cache(path)
        if header.embed_dim != clip_model.model.embed_dim:
            raise ValueError(f"{path}: embed_dim {header.embed_dim} does not match the loaded model")
        for key, want in (expect or {}).items():
            have = {"window_size": header.window_size, "window_stride": header.window_stride}.get(key, header.meta.get(key))
            if have is not None and have != want:
                raise ValueError(f"{path}: built with {key}={have!r}, current settings say {want!r}")
        t = torch.from_numpy(rows.view(np.int16) if rows.dtype == np.uint16 else rows)
        t = t.pin_memory().to(clip_model.model.device, non_blocking=True)
        if rows.dtype == np.uint16:
            t = t.view(torch.bfloat16)
        return cls(clip_model, t, ts, header.duration, header.meta)

    def save(self, path: str, block_rows: int = 1 << 16):
        import torch

        dtype = "bfloat16" if self.embeddings.dtype == torch.bfloat16 else "float32"
        with EmbeddingCacheWriter(path, self.embeddings.shape[1], dtype, True, self.meta, self.duration) as w:
            for lo in range(0, len(self.timestamps), block_rows):
                w.append(self.embeddings[lo:lo + block_rows], self.timestamps[lo:lo + block_rows])

    # ---------------------------------------------------------------------------------------- queries
    def __len__(self) -> int:
        return len(self.timestamps)

    def query_batch(self, queries: Union[str, Sequence[str]], top_k: int = None, threshold: float = None,
                    clip_duration: float = None) -> List[List[Dict]]:
        """One K4 pass for all queries.  Per query the same list of dicts as Phase1MVP.process_video
        (phase1_mvp.py:145-155) plus the clip interval of clip_extractor.py:175-183 ('start', 'end')."""
        import torch

        from ..utils.config import settings

        if isinstance(queries, str):
            queries = [queries]
        top_k = settings.TOP_K_RESULTS if top_k is None else top_k
        thr = settings.CONFIDENCE_THRESHOLD if threshold is None else threshold
        dur = settings.CLIP_DURATION if clip_duration is None else clip_duration
        if len(self) == 0:
            raise ValueError("No windows could be processed due to memory constraints")      # phase1_mvp.py:130-131
        txt = torch.from_numpy(self.clip_model.encode_text(list(queries))).to(self.model.device)
        k = int(top_k)
        if k <= 0:
            return [[] for _ in queries]
        scores, idx, iv, cnt = self.model.sim_topk(self.embeddings, txt, k, thr, self._ts_dev, index_base=0,
                                                   clip_duration=dur, video_duration=self.duration)
        scores, idx, iv, cnt = scores.cpu().numpy(), idx.cpu().numpy(), iv.cpu().numpy(), cnt.cpu().numpy()
        out = []
        for qi in range(len(queries)):
            res = []
            for r in range(int(cnt[qi])):
                i = int(idx[qi, r])
                res.append({"timestamp": self.timestamps[i], "confidence": float(scores[qi, r]), "phase": "phase1_mvp",
                            "window_index": i, "start": float(iv[qi, r, 0]), "end": float(iv[qi, r, 1])})
            out.append(res)
        return out

    def query(self, query: str, top_k: int = None, threshold: float = None) -> List[Dict]:
        return self.query_batch([query], top_k, threshold)[0]


def cache_path_for(video_path: str, cache_dir: str, model_name: str, fingerprint: str = "", sampling: str = "") -> str:
    """data/embeddings/<stem>-<digest>.b2emb; the digest covers path, size, mtime, model, weights and the sampling /
    window / resize settings (`sampling`: any string that changes when they do)."""
    st = os.stat(video_path)
    key = f"{os.path.abspath(video_path)}|{st.st_size}|{int(st.st_mtime)}|{model_name}|{fingerprint}|{sampling}"
    stem = os.path.splitext(os.path.basename(video_path))[0]
    return os.path.join(cache_dir, f"{stem}-{hashlib.sha256(key.encode()).hexdigest()[:12]}.b2emb")
