"""K1 alone on device-resident 1080p frames: RGB (3 B/px) vs the NV12 frame feed (1.5 B/px), CUDA-event times per call.
usage: python tools/bench_k1.py [frames=1024] [H=1080] [W=1920]"""
import json
import os
import sys

os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from b200clip import capi
from b200clip import open_clip as oc
from b200clip.model_configs import MODEL_CONFIGS
from b200clip.weights import random_state_dict
from bench import device_frames, peaks, rgb_to_nv12_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
dev = torch.device("cuda", 0)
cfg = MODEL_CONFIGS["ViT-B-32"]
model, _, _ = oc.create_model_and_transforms("ViT-B-32", state_dict=random_state_dict(cfg, 0), device=dev, max_images=8)
frames = device_frames(n, H, W, dev, 1)
nv = rgb_to_nv12_device(frames)
res = {}
for name, fn in (("rgb", lambda: model.preprocess_u8(frames, capi.RESIZE_REFERENCE)),
                 ("nv12", lambda: model.preprocess_nv12(nv, capi.RESIZE_REFERENCE))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    model.handle.profile_read(reset=True)
    model.handle.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 10
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    model.handle.profile_enable(False)
    prof = model.handle.profile_read(reset=True)
    ms = e0.elapsed_time(e1) / it
    res[name] = {"ms_per_call": round(ms, 4), "area_hpass_ms": round(prof["pre_area"]["ms"] / it, 4),
                 "vpass_ms": round(prof["pre_vpass"]["ms"] / it, 4),
                 "window_gbs": round(prof["pre_area"]["work"] / it / (prof["pre_area"]["ms"] / it / 1e3) / 1e9, 1),
                 "frac_of_hbm": round(prof["pre_area"]["work"] / it / (prof["pre_area"]["ms"] / it / 1e3) / 1e9 / peaks()["hbm_gbs"], 3)}
print(json.dumps({"frames": n, "hw": [H, W], **res}))
