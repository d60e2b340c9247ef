"""OpenCLIPModel: drop-in for /root/reference/src/models/openclip_model.py::OpenCLIPModel (:13-214) with the same
constructor, attributes (`model`, `preprocess`, `tokenizer`, `device`, `model_loaded`) and method semantics, backed
by libb200clip.so.

Differences that are deliberate (north_star: no CPU fallback):
  * the device is always the B200; `force_device="cpu"` raises instead of silently running on the host;
  * a failed load raises (the reference's @handle_model_loading_error swallows it and leaves model=None);
  * `encode_images` does not loop PIL over frames: the whole batch goes through the fused K1->K3 path with the
    transform-only resize mode (bit-identical pixels to PIL's transform, see tests/test_gpu_preprocess.py).
"""
from __future__ import annotations

from typing import List, Union

import numpy as np
import torch

from .. import capi
from .. import open_clip as b200_open_clip
from ..utils.config import settings
from ..utils.logger import get_logger

logger = get_logger(__name__)


class OpenCLIPModel:
    def __init__(self, force_device: str = None, state_dict=None, seed: int = 0):
        if force_device is not None and torch.device(force_device).type != "cuda":
            raise RuntimeError(f"OpenCLIPModel(force_device={force_device!r}): b200clip runs on the GPU only")
        if not torch.cuda.is_available():
            raise RuntimeError("b200clip needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(force_device) if force_device else torch.device("cuda", torch.cuda.current_device())
        self.model = None
        self.preprocess = None
        self.tokenizer = None
        self.model_loaded = False
        self._state_dict = state_dict
        self._seed = seed
        self.load_model()

    def load_model(self):
        """openclip_model.py:28-99 (create_model_and_transforms + get_tokenizer + eval)."""
        self.model, _, self.preprocess = b200_open_clip.create_model_and_transforms(
            settings.OPENCLIP_MODEL, pretrained=settings.OPENCLIP_PRETRAINED, device=self.device,
            state_dict=self._state_dict, seed=self._seed, max_images=settings.B200_MAX_IMAGES_PER_PASS)
        self.tokenizer = b200_open_clip.get_tokenizer(settings.OPENCLIP_MODEL)
        import os

        from ..tokenizer import HashTokenizer
        if isinstance(self.tokenizer, HashTokenizer) and self._state_dict is None and \
                isinstance(settings.OPENCLIP_PRETRAINED, str) and os.path.exists(settings.OPENCLIP_PRETRAINED):
            raise RuntimeError("a real checkpoint needs the real BPE vocabulary: set B200CLIP_BPE_VOCAB "
                               "(the stand-in HashTokenizer is for synthetic weights only)")
        self.model.eval()
        self.model_loaded = True
        logger.info(f"Loaded b200clip model: {settings.OPENCLIP_MODEL} on {self.device}")

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _to_uint8(images: np.ndarray) -> np.ndarray:
        # openclip_model.py:167-168,188-189
        if images.dtype != np.uint8:
            images = (images * 255).astype(np.uint8)
        return images

    def encode_images(self, images: Union[np.ndarray, List[np.ndarray]]) -> np.ndarray:
        """openclip_model.py:152-198: [N,H,W,3] -> float32 [N,E]; [H,W,3] -> float32 [1,E] (unit L2 norm)."""
        if isinstance(images, np.ndarray) and images.ndim == 4:
            batch = self._to_uint8(images)
        else:
            if isinstance(images, list):
                images = np.array(images)          # reference: a list falls into the single-image branch
            images = self._to_uint8(np.asarray(images))
            if images.ndim != 3:
                raise ValueError(f"expected one [H,W,3] image, got shape {images.shape}")
            batch = images[None]
        if batch.shape[0] == 0:
            return np.zeros((0, self.model.embed_dim), np.float32)
        return self.model.encode_frames_u8_host(batch, resize_mode=capi.RESIZE_BICUBIC, normalize=True)

    def encode_text(self, texts: Union[str, List[str]]) -> np.ndarray:
        """openclip_model.py:200-210."""
        if isinstance(texts, str):
            texts = [texts]
        tokens = self.tokenizer(texts)
        out = np.empty((tokens.shape[0], self.model.embed_dim), np.float32)
        tok = np.ascontiguousarray(tokens.numpy().astype(np.int64))
        self.model.handle.call("b200clip_encode_text_host", capi._p(tok), int(tok.shape[0]), capi._p(out), 1,
                               self.model._stream())
        return out

    def compute_similarity(self, image_embeddings: np.ndarray, text_embeddings: np.ndarray) -> np.ndarray:
        """openclip_model.py:212-214 (np.dot) on the GPU: float32 [N,Q]."""
        img = torch.from_numpy(np.ascontiguousarray(image_embeddings, dtype=np.float32)).to(self.device)
        txt = torch.from_numpy(np.ascontiguousarray(text_embeddings, dtype=np.float32)).to(self.device)
        return self.model.similarity(img, txt).cpu().numpy()
