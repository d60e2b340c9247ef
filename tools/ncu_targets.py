"""Launches each hot kernel of the step once after a warm-up, for `ncu --set full -k regex:<kernel>` captures
(tools/ncu_summary.py turns the reports into the text summaries under profiles/).

  python tools/ncu_targets.py b32      # 1024 x 1080p frames: K1 (RGB and NV12 feed), the tcgen05 GEMMs, persistent attention, head, K4
  python tools/ncu_targets.py l14attn  # ViT-L/14 attention, T = 257, 512 sequences x 16 heads (B200CLIP_ATTN_TC=1 selects the tcgen05 kernel)
"""
import ctypes
import os
import sys

os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from b200clip import capi
from b200clip import open_clip as oc
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config
from b200clip.weights import random_state_dict

what = sys.argv[1] if len(sys.argv) > 1 else "b32"
dev = torch.device("cuda", 0)
if what == "b32":
    sys.path.insert(0, ROOT)
    from bench import device_frames, rgb_to_nv12_device

    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    cfg = MODEL_CONFIGS["ViT-B-32"]
    model, _, _ = oc.create_model_and_transforms("ViT-B-32", state_dict=random_state_dict(cfg, 0), device=dev, max_images=n, max_texts=1)
    frames = device_frames(n, 1080, 1920, dev, 1)
    nv = rgb_to_nv12_device(frames)
    for _ in range(2):
        emb = model.encode_frames_u8(frames, capi.RESIZE_REFERENCE, normalize=True)
        emb2 = model.encode_frames_nv12(nv, capi.RESIZE_REFERENCE, normalize=True)
    torch.cuda.synchronize()
    print("b32 ok", float(emb.abs().sum()), float(emb2.abs().sum()))
else:
    h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-L-14"]), 0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n_seq, t, heads = 512, 257, 16
    qkv = (torch.randn(n_seq * t, 3 * heads * 64, device=dev) * 1.5).bfloat16()
    out = torch.empty(n_seq * t, heads * 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        h.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(48):         # two ViT-L/14 passes' worth of layers, back to back
        h.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, 0, st)
    e1.record()
    torch.cuda.synchronize()
    print("l14attn ok", float(out.float().abs().sum()), "ms per layer call (512 seq x 16 heads, T = 257):", round(e0.elapsed_time(e1) / 48, 4))
