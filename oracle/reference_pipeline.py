"""ORACLE (test infrastructure, not product): the reference's CPU pipeline for the hot path, restated end to end so
it can be TIMED on the GPU box's host cores (bench.py `cpu_baseline` and `--impl reference`) and used as the
checker in tests / smoke().  /root/reference does not exist on the GPU box, so this is a port ("kind": "port"):

  FrameExtractor (src/services/frame_extractor.py:89-91)  -> MemoryManager.resize_frame_for_memory
        cv2.resize(INTER_AREA) when OpenCV is installed (it is what the reference calls), else the numpy restatement
  OpenCLIPModel.encode_images (src/models/openclip_model.py:152-181): per image PIL -> open_clip transform
        (genuine torchvision/Pillow), torch.stack, batches of settings.BATCH_SIZE = 32, fp32 ViT forward
        (oracle/clip_ref.py), x / x.norm(), np.vstack
  OpenCLIPModel.encode_text / compute_similarity (:200-214), top-k + threshold (src/pipeline/phase1_mvp.py:145-155)
"""
from __future__ import annotations

import numpy as np
import torch

from . import clip_ref, phase1_ref, preprocess_ref
from .open_clip_shim import image_transform

BATCH_SIZE = 32  # src/utils/config.py:37


def resize_frame_for_memory(frame: np.ndarray, max_w: int = 512, max_h: int = 512) -> np.ndarray:
    h, w = frame.shape[:2]
    nw, nh = preprocess_ref.fit_size(w, h, max_w, max_h)
    if (nw, nh) == (w, h):
        return frame
    try:
        import cv2

        return cv2.resize(frame, (nw, nh), interpolation=cv2.INTER_AREA)
    except ImportError:
        return preprocess_ref.inter_area_resize(frame, nw, nh)


class ReferenceCPU:
    def __init__(self, model_name: str = "ViT-B-32", state_dict=None, seed: int = 0, threads: int | None = None):
        import os

        self.cfg = clip_ref.CONFIGS[model_name]
        sd = state_dict if state_dict is not None else clip_ref.init_state_dict(self.cfg, seed)
        self.model = clip_ref.CLIPRef(self.cfg, sd)
        self.preprocess = image_transform(self.cfg.image_size)
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)

    def encode_images(self, images: np.ndarray, shrink: bool = False) -> np.ndarray:
        from PIL import Image

        out = []
        for i in range(0, len(images), BATCH_SIZE):
            batch = images[i:i + BATCH_SIZE]
            tensors = []
            for img in batch:
                if img.dtype != np.uint8:
                    img = (img * 255).astype(np.uint8)
                if shrink:
                    img = resize_frame_for_memory(img)
                tensors.append(self.preprocess(Image.fromarray(img)))
            x = torch.stack(tensors)
            with torch.no_grad():
                e = self.model.encode_image(x)
                e = e / e.norm(dim=-1, keepdim=True)
            out.append(e.numpy())
        return np.vstack(out)

    def encode_text_tokens(self, tokens: torch.Tensor) -> np.ndarray:
        with torch.no_grad():
            t = self.model.encode_text(tokens)
            t = t / t.norm(dim=-1, keepdim=True)
        return t.numpy()

    def query(self, frames: np.ndarray, tokens: torch.Tensor, top_k: int, threshold: float, timestamps=None,
              shrink: bool = True):
        emb = self.encode_images(frames, shrink=shrink)
        txt = self.encode_text_tokens(tokens)
        sims = phase1_ref.compute_similarity(emb, txt)[:, 0]
        ts = timestamps if timestamps is not None else list(range(len(frames)))
        return phase1_ref.topk_threshold(sims, ts, top_k, threshold), sims
