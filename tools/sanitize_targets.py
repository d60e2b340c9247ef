"""Small invocations of the kernels added in round 2, for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
NV12 frame feed (fused + generic window conversion + host upload), K4 multi-pass top-k (k > 32) and its packed merge,
the tensor-core K4 path with the fp32 re-score, the wider head kernel, the persistent K1 launch form."""
import os
import sys

os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from b200clip import capi
from b200clip import distributed as D
from b200clip import open_clip as oc
from b200clip.model_configs import MODEL_CONFIGS
from b200clip.weights import random_state_dict

dev = torch.device("cuda", 0)
cfg = MODEL_CONFIGS["ViT-tiny-test"]
model, _, _ = oc.create_model_and_transforms("ViT-tiny-test", state_dict=random_state_dict(cfg, 0), device=dev, max_images=8)
rng = np.random.default_rng(0)
for h, w in [(1080, 1920), (720, 1280), (562, 1000)]:
    nv = rng.integers(0, 256, (3, h * 3 // 2, w), dtype=np.uint8)
    a = model.preprocess_nv12(torch.from_numpy(nv).cuda(), capi.RESIZE_REFERENCE)
    b = model.encode_frames_nv12_host(nv, capi.RESIZE_REFERENCE, True)
    rgb = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    c = model.encode_frames_u8_host(rgb, capi.RESIZE_REFERENCE, True)
torch.cuda.synchronize()
e = 64
img = torch.randn(5000, e, device=dev)
img = img / img.norm(dim=-1, keepdim=True)
txt = torch.randn(9, e, device=dev)
txt = txt / txt.norm(dim=-1, keepdim=True)
s, i, iv, c = model.sim_topk(img, txt[:2], 70, 0.0, torch.arange(5000, dtype=torch.float64), 0, 30.0, 5000.0)
s2, i2, iv2, c2 = model.sim_topk(img.bfloat16(), txt, 5, 0.0, None, 0, 30.0, 0.0)              # tensor-core path + re-score
q, k = 2, 70
msg = D.pack_message(s, i)
empty = D.pack_message(torch.full_like(s, float("-inf")), torch.full_like(i, -1))
gathered = torch.stack([empty, msg])
o = [torch.empty(q, k, device=dev), torch.empty(q, k, device=dev, dtype=torch.int64), torch.empty(q, k, 2, device=dev, dtype=torch.float64),
     torch.empty(q, device=dev, dtype=torch.int32)]
model.handle.call("b200clip_topk_merge_packed", capi._p(gathered), 2, q, k, 0.0, capi._p(None), 30.0, 0.0, capi._p(o[0]), capi._p(o[1]),
                  capi._p(o[2]), capi._p(o[3]), model._stream())
torch.cuda.synchronize()
assert torch.equal(o[1], i) and torch.equal(o[0], s)
emb = model.encode_frames_u8(torch.randint(0, 256, (37, 64, 64, 3), dtype=torch.uint8, device=dev))   # head kernel, 16 rows per CTA
torch.cuda.synchronize()
print("sanitize targets ok", float(emb.abs().sum()), int(c2.sum()))
