"""ORACLE (test infrastructure, not product): numpy restatement of the NV12 -> RGB conversion a video decoder's output
goes through before the reference's frame preprocessing.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Why it exists: the reference decodes with Decord / OpenCV VideoCapture (/root/reference/src/services/
frame_extractor.py:38-235), both of which hand RGB frames to Python after converting the decoder's native 4:2:0 output
on the CPU.  SURVEY.md section 8(f3) ("frame feed") moves that conversion onto the GPU so that 1.5 B/px instead of
3 B/px cross PCIe.  The conversion restated here is OpenCV's (third party, `opencv-python` of requirements.txt; the
installed 4.13 is the pin): `cv2.cvtColor(nv12, cv2.COLOR_YUV2RGB_NV12)` == `YUV420sp2RGB8Invoker` in
modules/imgproc/src/color_yuv.simd.hpp -- ITU-R BT.601 limited range, 20-bit fixed point, chroma sampled by
replication over each 2x2 block:

    y   = max(0, Y - 16) * 1220542
    R   = sat8((y + (1 << 19) + 1673527 * (V - 128)) >> 20)
    G   = sat8((y + (1 << 19) -  852492 * (V - 128) - 409993 * (U - 128)) >> 20)
    B   = sat8((y + (1 << 19) + 2116026 * (U - 128)) >> 20)

Pinning: tests/test_oracle_nv12.py checks `nv12_to_rgb` byte for byte against the installed cv2 on random and on
structured frames.  NV12 layout (as cv2 takes it): uint8 [H * 3 / 2, W]: rows [0, H) = Y, rows [H, 3H/2) = interleaved
U, V of the half-resolution chroma planes."""
from __future__ import annotations

import numpy as np

CY, CUB, CUG, CVG, CVR, SHIFT = 1220542, 2116026, -409993, -852492, 1673527, 20


def nv12_to_rgb(nv12: np.ndarray) -> np.ndarray:
    """uint8 [H*3/2, W] -> uint8 [H, W, 3] (RGB)."""
    hh, w = nv12.shape
    h = hh * 2 // 3
    if h % 2 or w % 2 or h * 3 // 2 != hh:
        raise ValueError(f"NV12 needs even width and height, got plane shape {nv12.shape}")
    y = np.maximum(0, nv12[:h].astype(np.int64) - 16) * CY + (1 << (SHIFT - 1))
    uv = nv12[h:].reshape(h // 2, w // 2, 2).astype(np.int64) - 128
    u = np.repeat(np.repeat(uv[..., 0], 2, 0), 2, 1)
    v = np.repeat(np.repeat(uv[..., 1], 2, 0), 2, 1)
    r = (y + CVR * v) >> SHIFT
    g = (y + CVG * v + CUG * u) >> SHIFT
    b = (y + CUB * u) >> SHIFT
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def rgb_to_nv12(rgb: np.ndarray) -> np.ndarray:
    """Test-input generator only (what an encoder + decoder would leave behind is not specified by the reference): BT.601
    limited-range luma / 2x2-averaged chroma in float, rounded.  uint8 [H, W, 3] -> uint8 [H*3/2, W]."""
    h, w, _ = rgb.shape
    if h % 2 or w % 2:
        raise ValueError("NV12 needs even width and height")
    f = rgb.astype(np.float64)
    r, g, b = f[..., 0], f[..., 1], f[..., 2]
    yy = 16 + (65.481 * r + 128.553 * g + 24.966 * b) / 255.0
    cb = 128 + (-37.797 * r - 74.203 * g + 112.0 * b) / 255.0
    cr = 128 + (112.0 * r - 93.786 * g - 18.214 * b) / 255.0
    out = np.empty((h * 3 // 2, w), np.uint8)
    out[:h] = np.clip(np.rint(yy), 0, 255)
    pool = lambda p: p.reshape(h // 2, 2, w // 2, 2).mean((1, 3))
    out[h:, 0::2] = np.clip(np.rint(pool(cb)), 0, 255)
    out[h:, 1::2] = np.clip(np.rint(pool(cr)), 0, 255)
    return out
