"""Device-resident embed throughput of one model geometry (used for BASELINE config 3: ViT-L/14 on 224x224 frames).
usage: python tools/bench_model.py --model ViT-L-14 --frames 512 [--hw 224 224]"""
import argparse
import json
import os
os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from b200clip import capi
from b200clip import open_clip as oc
from b200clip.model_configs import MODEL_CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ViT-L-14")
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--hw", type=int, nargs=2, default=[224, 224])
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
model, _, _ = oc.create_model_and_transforms(a.model, device="cuda:0", max_images=a.frames, max_texts=1, seed=0)
h = model.handle
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (a.frames, a.hw[0], a.hw[1], 3), device="cuda", dtype=torch.uint8, generator=g)
for _ in range(2):
    model.encode_frames_u8(frames, capi.RESIZE_REFERENCE)
torch.cuda.synchronize()
h.profile_read(reset=True)
h.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    model.encode_frames_u8(frames, capi.RESIZE_REFERENCE)
e1.record()
torch.cuda.synchronize()
h.profile_enable(False)
ms = e0.elapsed_time(e1) / a.steps
prof = h.profile_read(reset=True)
cfg = MODEL_CONFIGS[a.model]
line = {"model": a.model, "frames": a.frames, "hw": a.hw, "ms_per_pass": round(ms, 3), "frames_per_s": round(a.frames / ms * 1e3, 1),
        "kernels_ms_per_pass": {k: round(v["ms"] / a.steps, 3) for k, v in prof.items() if v["launches"]}}
gm = prof["gemm"]
if gm["ms"] > 0:
    line["gemm_tflops"] = round(gm["work"] / (gm["ms"] / 1e3) / 1e12, 1)
print(json.dumps(line))
