"""The `open_clip` module surface the reference binds to, served by libb200clip.so.

The reference (src/models/openclip_model.py) uses exactly this slice of open_clip:
    model, _, preprocess = open_clip.create_model_and_transforms(name, pretrained=..., device=...)   (:77-81)
    tokenizer = open_clip.get_tokenizer(name)                                                          (:82)
    model.eval(); model.encode_image(x[B,3,S,S]); model.encode_text(tokens[Q,77])                     (:83,177,205-208)
    preprocess(PIL.Image) -> FloatTensor[3,S,S]                                                         (:171,193)
so `sys.modules["open_clip"] = b200clip.open_clip` makes the unmodified reference wrapper run on the B200
kernels.  The model object additionally offers the fused fast paths
(`encode_frames_u8`, `encode_frames_u8_host`, `sim_topk`) that skip PIL and fp32 CHW tensors entirely.

PyTorch is used here for device memory and streams only; every FLOP runs in the C-ABI library.
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi
from .model_configs import MODEL_CONFIGS, ModelConfig, to_capi_config
from .tokenizer import get_tokenizer  # noqa: F401  (re-exported: open_clip.get_tokenizer)
from .weights import load_checkpoint, random_state_dict


def _device_index(device) -> int:
    if device is None:
        return torch.cuda.current_device() if torch.cuda.is_available() else 0
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError(f"b200clip has no {d.type} path: the model runs on a B200 (cuda) device only")
    return d.index if d.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)


class B200CLIP:
    """CLIP (ViT image tower + text tower) whose forward is libb200clip.so."""

    def __init__(self, cfg: ModelConfig, state_dict, device=None, max_images: int = 0, max_texts: int = 0):
        self.cfg = cfg
        self.device_index = _device_index(device)
        self.device = torch.device("cuda", self.device_index)
        self.handle = capi.Handle(to_capi_config(cfg), self.device_index)
        self.handle.load_state_dict(state_dict)
        if max_images or max_texts:
            self.handle.reserve(max_images, max_texts)
        self.embed_dim = cfg.embed_dim
        self._comm_cache = {}       # (group, device) -> ncclComm_t of torch's process group (0 = not NCCL)
        self._msg_cache = {}        # (q, k, world) -> (message, gathered) buffers of the generic exchange

    # ---- open_clip surface -------------------------------------------------------------------
    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def _stream(self):
        return capi.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """model.encode_image(x[B,3,S,S]) -> [B,E] fp32 (un-normalised unless `normalize`)."""
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] != self.cfg.image_size \
                or image.shape[3] != self.cfg.image_size:
            raise ValueError(f"expected [B,3,{self.cfg.image_size},{self.cfg.image_size}], got {tuple(image.shape)}")
        x = image.to(self.device, torch.float32).contiguous()
        out = torch.empty(x.shape[0], self.embed_dim, device=self.device, dtype=torch.float32)
        self.handle.call("b200clip_encode_image_chw", capi._p(x), int(x.shape[0]), capi._p(out), capi.F32,
                         int(normalize), self._stream())
        return out

    def encode_text(self, text: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """model.encode_text(tokens[Q,ctx]) -> [Q,E] fp32."""
        if text.dim() != 2 or text.shape[1] != self.cfg.text_ctx:
            raise ValueError(f"expected [Q,{self.cfg.text_ctx}] token ids, got {tuple(text.shape)}")
        t = text.to(self.device, torch.int64).contiguous()
        out = torch.empty(t.shape[0], self.embed_dim, device=self.device, dtype=torch.float32)
        self.handle.call("b200clip_encode_text", capi._p(t), int(t.shape[0]), capi._p(out), int(normalize),
                         self._stream())
        return out

    # ---- fused fast paths ----------------------------------------------------------------------
    def preprocess_u8(self, frames: torch.Tensor, resize_mode: int = capi.RESIZE_REFERENCE,
                      chw: bool = False) -> torch.Tensor:
        """uint8 [N,H,W,3] (cuda) -> bf16 patch rows [N*g*g, patch_k] or fp32 [N,3,S,S]."""
        f = self._check_frames(frames)
        n, h, w = int(f.shape[0]), int(f.shape[1]), int(f.shape[2])
        if chw:
            out = torch.empty(n, 3, self.cfg.image_size, self.cfg.image_size, device=self.device, dtype=torch.float32)
            self.handle.call("b200clip_preprocess_u8_chw", capi._p(f), n, h, w, h * w * 3, w * 3, resize_mode,
                             capi._p(out), self._stream())
        else:
            g = self.cfg.image_size // self.cfg.patch
            out = torch.empty(n * g * g, self.cfg.patch_k, device=self.device, dtype=torch.bfloat16)
            self.handle.call("b200clip_preprocess_u8", capi._p(f), n, h, w, h * w * 3, w * 3, resize_mode,
                             capi._p(out), self._stream())
        return out

    def _check_frames(self, frames: torch.Tensor) -> torch.Tensor:
        if frames.dim() != 4 or frames.shape[3] != 3 or frames.dtype != torch.uint8:
            raise ValueError(f"expected uint8 [N,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
        if frames.device.type != "cuda":
            raise ValueError("frames must live on the GPU for this call (use encode_frames_u8_host for host data)")
        return frames.contiguous()

    def encode_frames_u8(self, frames: torch.Tensor, resize_mode: int = capi.RESIZE_REFERENCE,
                         normalize: bool = True, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Device-resident decoded frames uint8 [N,H,W,3] -> embeddings [N,E] (preprocess + ViT + head fused)."""
        f = self._check_frames(frames)
        n, h, w = int(f.shape[0]), int(f.shape[1]), int(f.shape[2])
        out = torch.empty(n, self.embed_dim, device=self.device, dtype=out_dtype)
        dt = capi.BF16 if out_dtype == torch.bfloat16 else capi.F32
        self.handle.call("b200clip_encode_frames_u8", capi._p(f), n, h, w, h * w * 3, w * 3, resize_mode, capi._p(out),
                         dt, int(normalize), self._stream())
        return out

    def encode_frames_u8_host(self, frames, resize_mode: int = capi.RESIZE_REFERENCE, normalize: bool = True,
                              out: np.ndarray | None = None) -> np.ndarray:
        """HOST frames (numpy uint8 [N,H,W,3] or a CPU tensor, pinned or pageable) -> HOST float32 [N,E].
        H2D copies are double buffered against compute inside the library; returns when `out` is complete."""
        if isinstance(frames, torch.Tensor):
            if frames.device.type != "cpu" or frames.dtype != torch.uint8:
                raise ValueError("expected a CPU uint8 tensor")
            arr = frames.contiguous()
            shape = tuple(arr.shape)
        else:
            arr = np.ascontiguousarray(frames)
            if arr.dtype != np.uint8:
                raise ValueError("expected uint8 frames")
            shape = arr.shape
        if len(shape) != 4 or shape[3] != 3:
            raise ValueError(f"expected [N,H,W,3], got {shape}")
        n, h, w = int(shape[0]), int(shape[1]), int(shape[2])
        if out is None:
            out = np.empty((n, self.embed_dim), np.float32)
        self.handle.call("b200clip_encode_frames_u8_host", capi._p(arr), n, h, w, resize_mode, capi._p(out),
                         int(normalize), self._stream())
        return out

    # ---- NV12 frame feed (decoder output; cv2 layout uint8 [N, H*3/2, W]: Y rows then interleaved UV rows) ----------
    def _check_nv12(self, frames, device_type: str):
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dim() != 3 or frames.dtype != torch.uint8 or frames.shape[1] % 3 or frames.shape[2] % 2 or \
                (frames.shape[1] * 2 // 3) % 2:
            raise ValueError(f"expected NV12 uint8 [N, H*3/2, W] with even H and W, got {frames.dtype} {tuple(frames.shape)}")
        if frames.device.type != device_type:
            raise ValueError(f"NV12 frames must live on the {device_type} for this call")
        return frames.contiguous()

    def preprocess_nv12(self, frames: torch.Tensor, resize_mode: int = capi.RESIZE_REFERENCE,
                        chw: bool = False) -> torch.Tensor:
        """NV12 uint8 [N, H*3/2, W] (cuda) -> bf16 patch rows [N*g*g, patch_k] or fp32 [N,3,S,S]; pixels identical to
        cv2.cvtColor(COLOR_YUV2RGB_NV12) + preprocess_u8."""
        f = self._check_nv12(frames, "cuda")
        n, hh, w = (int(v) for v in f.shape)
        h = hh * 2 // 3
        y_ptr = f.data_ptr()
        if chw:
            out = torch.empty(n, 3, self.cfg.image_size, self.cfg.image_size, device=self.device, dtype=torch.float32)
            po, co = None, out
        else:
            g = self.cfg.image_size // self.cfg.patch
            out = torch.empty(n * g * g, self.cfg.patch_k, device=self.device, dtype=torch.bfloat16)
            po, co = out, None
        self.handle.call("b200clip_preprocess_nv12", capi._p(y_ptr), capi._p(y_ptr + h * w), n, h, w, hh * w, hh * w, w,
                         resize_mode, capi._p(po), capi._p(co), self._stream())
        return out

    def encode_frames_nv12(self, frames: torch.Tensor, resize_mode: int = capi.RESIZE_REFERENCE, normalize: bool = True,
                           out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Device-resident NV12 frames [N, H*3/2, W] -> embeddings [N,E] (colour conversion fused into K1)."""
        f = self._check_nv12(frames, "cuda")
        n, hh, w = (int(v) for v in f.shape)
        h = hh * 2 // 3
        out = torch.empty(n, self.embed_dim, device=self.device, dtype=out_dtype)
        dt = capi.BF16 if out_dtype == torch.bfloat16 else capi.F32
        y_ptr = f.data_ptr()
        self.handle.call("b200clip_encode_frames_nv12", capi._p(y_ptr), capi._p(y_ptr + h * w), n, h, w, hh * w, hh * w, w,
                         resize_mode, capi._p(out), dt, int(normalize), self._stream())
        return out

    def encode_frames_nv12_host(self, frames, resize_mode: int = capi.RESIZE_REFERENCE, normalize: bool = True,
                                out=None):
        """HOST NV12 frames (numpy / CPU tensor uint8 [N, H*3/2, W]) -> float32 [N,E] (host numpy, or the device tensor
        passed as `out`): half the PCIe bytes of encode_frames_u8_host."""
        f = self._check_nv12(frames, "cpu")
        n, hh, w = (int(v) for v in f.shape)
        if out is None:
            out = np.empty((n, self.embed_dim), np.float32)
        self.handle.call("b200clip_encode_frames_nv12_host", capi._p(f), n, hh * 2 // 3, w, resize_mode, capi._p(out),
                         int(normalize), self._stream())
        return out

    def similarity(self, img_emb: torch.Tensor, txt_emb: torch.Tensor) -> torch.Tensor:
        """compute_similarity: [N,E] x [Q,E] -> fp32 [N,Q]."""
        img = img_emb.contiguous()
        txt = txt_emb.to(self.device, torch.float32).contiguous()
        out = torch.empty(img.shape[0], txt.shape[0], device=self.device, dtype=torch.float32)
        dt = capi.BF16 if img.dtype == torch.bfloat16 else capi.F32
        self.handle.call("b200clip_similarity", capi._p(img), dt, int(img.shape[0]), int(img.shape[1]), capi._p(txt),
                         int(txt.shape[0]), capi._p(out), self._stream())
        return out

    def sim_topk(self, img_emb: torch.Tensor, txt_emb: torch.Tensor, k: int, threshold: float = -float("inf"),
                 timestamps: torch.Tensor | None = None, index_base: int = 0, clip_duration: float = 30.0,
                 video_duration: float = 0.0):
        """Fused similarity + top-k + threshold + clip intervals (K4).  Returns device tensors
        (scores [Q,k] f32, idx [Q,k] i64, intervals [Q,k,2] f64, counts [Q] i32)."""
        img = img_emb.contiguous()
        if img.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("embeddings must be float32 or bfloat16")
        txt = txt_emb.to(self.device, torch.float32).contiguous()
        q = int(txt.shape[0])
        scores = torch.empty(q, k, device=self.device, dtype=torch.float32)
        idx = torch.empty(q, k, device=self.device, dtype=torch.int64)
        iv = torch.empty(q, k, 2, device=self.device, dtype=torch.float64)
        cnt = torch.empty(q, device=self.device, dtype=torch.int32)
        ts = None
        if timestamps is not None:
            ts = timestamps.to(self.device, torch.float64).contiguous()
        dt = capi.BF16 if img.dtype == torch.bfloat16 else capi.F32
        thr = float(max(threshold, -3.0e38))
        self.handle.call("b200clip_sim_topk", capi._p(img), dt, int(img.shape[0]), int(img.shape[1]), capi._p(txt), q,
                         int(k), thr, capi._p(ts), int(index_base), float(clip_duration), float(video_duration),
                         capi._p(scores), capi._p(idx), capi._p(iv), capi._p(cnt), self._stream())
        return scores, idx, iv, cnt

    def sim_topk_sharded(self, img_emb: torch.Tensor, txt_emb: torch.Tensor, k: int, threshold: float = -float("inf"),
                         timestamps: torch.Tensor | None = None, index_base: int = 0, clip_duration: float = 30.0,
                         video_duration: float = 0.0, group=None):
        """K4 over a row-sharded embedding matrix (this rank holds rows [index_base, index_base + len(img_emb))):
        local top-k -> ONE all-gather of the packed candidate message -> the same merge on every rank.  Same outputs
        as sim_topk, identical on all ranks and bit-identical to sim_topk over the unsharded matrix.  On NCCL the whole
        exchange runs inside libb200clip.so (b200clip_sim_topk_nccl: no torch kernel, no host sync); other backends
        move the message with torch.distributed."""
        from . import distributed as D

        rank, world = D.world_info(group)
        if world == 1:
            return self.sim_topk(img_emb, txt_emb, k, threshold, timestamps, index_base, clip_duration, video_duration)
        img = img_emb.contiguous()
        txt = txt_emb.to(self.device, torch.float32).contiguous()
        q = int(txt.shape[0])
        scores = torch.empty(q, k, device=self.device, dtype=torch.float32)
        idx = torch.empty(q, k, device=self.device, dtype=torch.int64)
        iv = torch.empty(q, k, 2, device=self.device, dtype=torch.float64)
        cnt = torch.empty(q, device=self.device, dtype=torch.int32)
        ts = timestamps.to(self.device, torch.float64).contiguous() if timestamps is not None else None
        dt = capi.BF16 if img.dtype == torch.bfloat16 else capi.F32
        thr = float(max(threshold, -3.0e38))
        key = (id(group), self.device_index)
        if key not in self._comm_cache:
            self._comm_cache[key] = D.nccl_comm_ptr(self.device, group)
        comm = self._comm_cache[key]
        if comm:
            self.handle.call("b200clip_sim_topk_nccl", capi._p(comm), rank, world, capi._p(img), dt, int(img.shape[0]),
                             int(img.shape[1]), capi._p(txt), q, int(k), thr, capi._p(ts), int(index_base),
                             float(clip_duration), float(video_duration), capi._p(scores), capi._p(idx), capi._p(iv),
                             capi._p(cnt), self._stream())
            return scores, idx, iv, cnt
        # generic transport: K4 writes into a message buffer, torch.distributed gathers it, the merge kernel reads the
        # gathered messages in place
        mb = D.msg_bytes(q, k)
        bkey = (q, k, world)
        if bkey not in self._msg_cache:
            self._msg_cache[bkey] = (torch.zeros(mb, dtype=torch.uint8, device=self.device),
                                     torch.empty(world, mb, dtype=torch.uint8, device=self.device))
        msg, gathered = self._msg_cache[bkey]
        self.handle.call("b200clip_sim_topk", capi._p(img), dt, int(img.shape[0]), int(img.shape[1]), capi._p(txt), q, int(k),
                         -3.0e38, capi._p(None), int(index_base), 0.0, 0.0, capi._p(msg.data_ptr() + q * k * 8),
                         capi._p(msg.data_ptr()), capi._p(None), capi._p(None), self._stream())
        if dist_backend_is_cpu_only(group):
            gathered.copy_(D.exchange_messages(msg.cpu(), group))
        else:
            D.exchange_messages(msg, group, out=gathered)
        self.handle.call("b200clip_topk_merge_packed", capi._p(gathered), world, q, int(k), thr, capi._p(ts),
                         float(clip_duration), float(video_duration), capi._p(scores), capi._p(idx), capi._p(iv),
                         capi._p(cnt), self._stream())
        return scores, idx, iv, cnt

    def topk_merge(self, cand_scores: torch.Tensor, cand_idx: torch.Tensor, threshold: float = -float("inf"),
                   timestamps: torch.Tensor | None = None, clip_duration: float = 30.0, video_duration: float = 0.0):
        """Merge [g,Q,k] candidate lists (global indices) into the global top-k."""
        cs = cand_scores.to(self.device, torch.float32).contiguous()
        ci = cand_idx.to(self.device, torch.int64).contiguous()
        g, q, k = (int(v) for v in cs.shape)
        scores = torch.empty(q, k, device=self.device, dtype=torch.float32)
        idx = torch.empty(q, k, device=self.device, dtype=torch.int64)
        iv = torch.empty(q, k, 2, device=self.device, dtype=torch.float64)
        cnt = torch.empty(q, device=self.device, dtype=torch.int32)
        ts = timestamps.to(self.device, torch.float64).contiguous() if timestamps is not None else None
        thr = float(max(threshold, -3.0e38))
        self.handle.call("b200clip_topk_merge", capi._p(cs), capi._p(ci), g, q, k, thr, capi._p(ts),
                         float(clip_duration), float(video_duration), capi._p(scores), capi._p(idx), capi._p(iv),
                         capi._p(cnt), self._stream())
        return scores, idx, iv, cnt


def dist_backend_is_cpu_only(group=None) -> bool:
    import torch.distributed as dist

    return dist.get_backend(group) == "gloo"


class _Preprocess:
    """open_clip's eval `image_transform` (PIL.Image -> FloatTensor[3,S,S]) evaluated by the K1 kernel in
    transform-only mode (Pillow-bicubic(aa) Resize -> CenterCrop -> ToTensor -> Normalize), bit-exact."""

    def __init__(self, model: B200CLIP):
        self.model = model

    def __call__(self, img) -> torch.Tensor:
        arr = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img)
        if arr.ndim != 3 or arr.shape[2] != 3 or arr.dtype != np.uint8:
            raise ValueError("preprocess expects an RGB uint8 image")
        f = torch.from_numpy(np.ascontiguousarray(arr)).to(self.model.device).unsqueeze(0)
        return self.model.preprocess_u8(f, capi.RESIZE_BICUBIC, chw=True)[0]


def create_model_and_transforms(model_name: str, pretrained=None, device=None, state_dict=None, seed: int = 0,
                                max_images: int = 0, max_texts: int = 0, **_kw):
    """open_clip.create_model_and_transforms look-alike -> (model, preprocess_train(None), preprocess_val).

    Weights: `state_dict` (open_clip key names) if given; else `pretrained` may be a path to a checkpoint in
    open_clip's native layout.  A tag such as "openai" cannot be downloaded here (no network): that RAISES unless
    B200CLIP_ALLOW_SYNTHETIC=1 opts into a seeded random init (tests / bench) -- never a silent stand-in.

    Activation (open_clip semantics): QuickGELU for `pretrained="openai"` and for "*-quickgelu" model names, erf GELU
    for every other tag / checkpoint; `quick_gelu=True/False` overrides.
    """
    import dataclasses
    import os

    from .tokenizer import allow_synthetic

    name = model_name.replace("/", "-")
    named_quick = name.lower().endswith("-quickgelu")
    key = name[:-len("-quickgelu")] if named_quick else name
    if key not in MODEL_CONFIGS:
        raise ValueError(f"unknown model '{model_name}'; available: {sorted(MODEL_CONFIGS)}")
    cfg = MODEL_CONFIGS[key]
    quick = _kw.get("quick_gelu", _kw.get("force_quick_gelu"))
    if not quick and "quick_gelu" not in _kw:
        quick = named_quick or pretrained is None or str(pretrained).lower() == "openai"
    if bool(quick) != cfg.quick_gelu:
        cfg = dataclasses.replace(cfg, quick_gelu=bool(quick))
    if state_dict is None:
        if isinstance(pretrained, str) and os.path.exists(pretrained):
            state_dict = load_checkpoint(pretrained)
        elif allow_synthetic():
            import logging

            logging.getLogger(__name__).warning(
                "no checkpoint for pretrained=%r is available offline; B200CLIP_ALLOW_SYNTHETIC=1 -> seeded random "
                "init (seed=%d)", pretrained, seed)
            state_dict = random_state_dict(cfg, seed)
        else:
            raise RuntimeError(
                f"no checkpoint for pretrained={pretrained!r}: pass a path to an open_clip state dict (or state_dict=), "
                "or set B200CLIP_ALLOW_SYNTHETIC=1 to run seeded random weights (results are then meaningless)")
    model = B200CLIP(cfg, state_dict, device=device, max_images=max_images, max_texts=max_texts)
    pre = _Preprocess(model)
    return model, pre, pre
