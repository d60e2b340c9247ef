"""BASELINE config 4: 256 text queries x 1M cached frame embeddings (bf16, E=512), top_k=5: fused similarity GEMM +
top-k (sim_topk_tc_kernel) vs the HBM-streaming kernel (B200CLIP_SIM_SIMT=1), with the roofline numbers of
SURVEY.md section 8(d)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from b200clip import capi
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config

n, q, k, e = 1_000_000, 256, 5, 512
h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
g = torch.Generator(device="cuda").manual_seed(0)
img = torch.randn(n, e, device="cuda", generator=g)
img = (img / img.norm(dim=-1, keepdim=True)).bfloat16()
txt = torch.randn(q, e, device="cuda", generator=g)
txt = txt / txt.norm(dim=-1, keepdim=True)
s = torch.empty(q, k, device="cuda")
i = torch.empty(q, k, device="cuda", dtype=torch.int64)
iv = torch.empty(q, k, 2, device="cuda", dtype=torch.float64)
c = torch.empty(q, device="cuda", dtype=torch.int32)
st = capi.c_void_p(torch.cuda.current_stream().cuda_stream)


def call():
    h.call("b200clip_sim_topk", capi._p(img), capi.BF16, n, e, capi._p(txt), q, k, 0.1, None, 0, 30.0, 0.0, capi._p(s),
           capi._p(i), capi._p(iv), capi._p(c), st)


for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for _ in range(iters):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
ref = torch.topk(img.float() @ txt.bfloat16().float().t(), k, dim=0)
bytes_ = n * e * 2 + q * e * 4 + q * k * 12
print(json.dumps({"config": "256 queries x 1M bf16 embeddings (E=512), top_k=5",
                  "kernel": "hbm-streaming (simt)" if os.environ.get("B200CLIP_SIM_SIMT") else "tcgen05 fused top-k",
                  "ms": ms, "tflops": 2.0 * n * q * e / ms / 1e9, "hbm_gbs_algorithmic": bytes_ / ms / 1e6,
                  "top1_matches_torch": float((i[:, 0] == ref.indices[0]).float().mean())}))
