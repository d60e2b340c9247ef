"""Seeded synthetic inputs shared by the golden generator, the parity tests, smoke() and bench.py.

Structured frames (random solid background + random filled rectangles + a low-frequency gradient) give
image embeddings that are far from collinear, unlike uniform noise."""
from __future__ import annotations

import numpy as np


def structured_frames(n: int, h: int = 224, w: int = 224, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        bg = rng.integers(0, 256, 3).astype(np.float32)
        gx, gy = rng.uniform(-0.5, 0.5, 2)
        img = bg[None, None, :] + (gx * (xx - w / 2) * (255.0 / w) + gy * (yy - h / 2) * (255.0 / h))[..., None]
        for _ in range(int(rng.integers(1, 6))):
            x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
            x1, y1 = int(rng.integers(x0 + 4, w)), int(rng.integers(y0 + 4, h))
            img[y0:y1, x0:x1, :] = rng.integers(0, 256, 3).astype(np.float32)
        img += rng.normal(0, 3.0, img.shape).astype(np.float32)
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def noise_frames(n: int, h: int, w: int, seed: int = 99) -> np.ndarray:
    """Uniform noise: the adversarial input for the resize kernels (every tap matters)."""
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 3), dtype=np.uint8)


QUERIES = ["a person walks across the street", "red car drives fast", "the dog jumps over a fence"]


def write_test_video(path, n: int, w: int = 96, h: int = 64, fps: float = 8.0) -> str:
    """A small deterministic mp4 (OpenCV's mp4v encoder): per-frame background level, a sliding gradient and a moving
    block, so that every frame decodes to distinct pixels."""
    import cv2

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


def write_frames_video(path, frames: np.ndarray, fps: float = 8.0) -> str:
    """RGB uint8 frames [N,H,W,3] -> mp4 (OpenCV's mp4v encoder, deterministic for a given OpenCV build)."""
    import cv2

    h, w = int(frames.shape[1]), int(frames.shape[2])
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    assert vw.isOpened()
    for f in frames:
        vw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    vw.release()
    return str(path)
