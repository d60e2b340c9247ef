"""ctypes binding of libb200clip.so (include/b200clip.h).  No torch types cross this boundary: only raw
pointers (`tensor.data_ptr()`, `ndarray.ctypes.data`) and sizes.

There is deliberately no fallback: if the shared library is missing or the device is not sm_100 every entry
point raises.  The oracle (oracle/) is never imported from here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200clip.so")

RESIZE_REFERENCE = 0
RESIZE_BILINEAR_AA = 1
RESIZE_BICUBIC = 2
INPUT_BGR = 0x100     # OR into resize_mode of the uint8 entry points: frames in OpenCV's BGR order (B200CLIP_INPUT_BGR)
F32 = 0
BF16 = 1

ERRORS = {-1: "E_ARG", -2: "E_SHAPE", -3: "E_CUDA", -4: "E_ARCH", -5: "E_STATE", -6: "E_NOMEM", -7: "E_NCCL"}


class B200ClipError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200clip {ERRORS.get(code, code)}: {msg}")
        self.code = code


class Config(Structure):
    _fields_ = [
        ("image_size", c_int32), ("patch", c_int32), ("width", c_int32), ("layers", c_int32),
        ("heads", c_int32), ("mlp_dim", c_int32), ("embed_dim", c_int32), ("act", c_int32),
        ("ln_eps", c_float),
        ("text_ctx", c_int32), ("text_vocab", c_int32), ("text_width", c_int32), ("text_heads", c_int32),
        ("text_layers", c_int32), ("text_mlp_dim", c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/b200clip.h declares (tests check this)
SIGNATURES = {
    "b200clip_version": (c_char_p, []),
    "b200clip_last_error": (c_char_p, [c_void_p]),
    "b200clip_create": (c_int, [POINTER(Config), c_int, POINTER(c_void_p)]),
    "b200clip_destroy": (c_int, [c_void_p]),
    "b200clip_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "b200clip_finalize": (c_int, [c_void_p]),
    "b200clip_reserve": (c_int, [c_void_p, c_int, c_int]),
    "b200clip_preprocess_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p,
                                       c_void_p]),
    "b200clip_preprocess_u8_chw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p,
                                           c_void_p]),
    "b200clip_encode_patches": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b200clip_encode_image_chw": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b200clip_encode_frames_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p,
                                          c_int, c_int, c_void_p]),
    "b200clip_encode_frames_u8_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                               c_void_p]),
    "b200clip_preprocess_nv12": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int,
                                         c_void_p, c_void_p, c_void_p]),
    "b200clip_encode_frames_nv12": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int64,
                                            c_int, c_void_p, c_int, c_int, c_void_p]),
    "b200clip_encode_frames_nv12_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                                 c_void_p]),
    "b200clip_encode_text": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "b200clip_encode_text_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "b200clip_sim_topk": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int, c_float, c_void_p,
                                  c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200clip_sim_topk_dense": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int, c_float, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200clip_similarity": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "b200clip_topk_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_double,
                                    c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200clip_topk_msg_bytes": (c_int64, [c_int, c_int]),
    "b200clip_sim_topk_nccl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int,
                                       c_float, c_void_p, c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p]),
    "b200clip_topk_merge_nccl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p,
                                         c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200clip_topk_merge_packed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_double, c_double,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200clip_gemm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_int, c_void_p]),
    "b200clip_layernorm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float,
                                        c_void_p]),
    "b200clip_attention_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200clip_profile_enable": (c_int, [c_void_p, c_int]),
    "b200clip_profile_read": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_double), POINTER(c_int64), c_int]),
    "b200clip_launch_count": (c_int64, [c_void_p]),
    "b200clip_reset_launch_count": (None, [c_void_p]),
    "b200clip_transfer_bytes": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
}

_lib = None


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Loads libb200clip.so (built in-tree by __graft_entry__.build() / `make -C csrc`).  Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("B200CLIP_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _p(x) -> c_void_p:
    """Raw pointer of a torch tensor / numpy array / int / None."""
    if x is None:
        return c_void_p(0)
    if isinstance(x, int):
        return c_void_p(x)
    if hasattr(x, "data_ptr"):
        return c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):
        return c_void_p(x.ctypes.data)
    raise TypeError(f"cannot take a pointer of {type(x)}")


class Handle:
    """Owns one b200clip_handle (one per GPU / rank)."""

    def __init__(self, cfg: Config, device: int = 0):
        self.lib = load_library()
        self._h = c_void_p()
        self.cfg = cfg
        self.device = device
        rc = self.lib.b200clip_create(byref(cfg), device, byref(self._h))
        if rc != 0:
            raise B200ClipError(rc, (self.lib.b200clip_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.b200clip_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise B200ClipError(rc, (self.lib.b200clip_last_error(self._h) or b"").decode())

    def call(self, name: str, *args):
        self._check(getattr(self.lib, name)(self._h, *args))

    # ---- weights
    def set_weight(self, name: str, array):
        import numpy as np

        a = np.ascontiguousarray(array, dtype=np.float32)
        shape = (c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
        self.call("b200clip_set_weight", name.encode(), _p(a), shape, a.ndim)

    def load_state_dict(self, sd):
        for k, v in sd.items():
            if k == "logit_scale":
                continue
            self.set_weight(k, v.detach().cpu().numpy() if hasattr(v, "detach") else v)
        self.call("b200clip_finalize")

    def reserve(self, max_images: int, max_texts: int = 0):
        self.call("b200clip_reserve", int(max_images), int(max_texts))

    PROFILE_CLASSES = ("gemm", "attention", "layernorm", "preprocess", "head", "sim_topk", "misc", "pre_area", "pre_hpass",
                       "pre_vpass", "comm", "gemm_small")

    def profile_enable(self, on: bool = True):
        self.call("b200clip_profile_enable", int(on))

    def profile_read(self, reset: bool = True) -> dict:
        """{class: {"ms", "work", "launches"}} summed over the timed launches since the last reset."""
        out = {}
        for i, name in enumerate(self.PROFILE_CLASSES):
            ms, work, n = c_double(), c_double(), c_int64()
            last = i == len(self.PROFILE_CLASSES) - 1
            self.call("b200clip_profile_read", i, byref(ms), byref(work), byref(n), int(reset and last))
            out[name] = {"ms": ms.value, "work": work.value, "launches": n.value}
        return out

    @property
    def launches(self) -> int:
        return int(self.lib.b200clip_launch_count(self._h))

    def reset_launches(self):
        self.lib.b200clip_reset_launch_count(self._h)

    def transfer_bytes(self, reset: bool = True) -> tuple:
        """(host->device, device->host) bytes moved by the *_host entry points since the last reset."""
        a, b = c_int64(), c_int64()
        self.call("b200clip_transfer_bytes", byref(a), byref(b), int(reset))
        return a.value, b.value
