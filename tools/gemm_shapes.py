"""Per-shape timing of the ViT-B/32 GEMMs (M = 3600*50 token rows) for the 1-CTA and 2-CTA kernels.
usage: python tools/gemm_shapes.py [1cta]   (B200CLIP_GEMM_1CTA is set for the 1cta variant)"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "1cta":
    os.environ["B200CLIP_GEMM_1CTA"] = "1"
import torch

from b200clip import capi
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config

h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
M = int(os.environ.get("GEMM_M", 180000))
shapes = [("patch", M * 49 // 50, 768, 3072, 0, 0, 0), ("qkv", M, 2304, 768, 1, 0, 0), ("out", M, 768, 768, 1, 1, 0),
          ("fc", M, 3072, 768, 1, 0, 1), ("proj", M, 768, 3072, 1, 1, 0), ("fc_noact", M, 3072, 768, 1, 0, 0),
          ("qkv_nobias", M, 2304, 768, 0, 0, 0)]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {}
for name, m, n, k, bias, resid, act in shapes:
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    b = torch.randn(n, device="cuda") if bias else None
    out = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)

    def call():
        h.call("b200clip_gemm_bf16", capi._p(a), capi._p(w), capi._p(out), m, n, k, capi._p(b),
               capi._p(out if resid else None), act, st)

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 10
    e0.record()
    for _ in range(it):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    for _ in range(3):
        torch.matmul(a, w.t())
    e0.record()
    for _ in range(it):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    cms = e0.elapsed_time(e1) / it
    res[name] = {"ms": round(ms, 4), "tflops": round(2.0 * m * n * k / ms / 1e9, 1), "cublas_tflops": round(2.0 * m * n * k / cms / 1e9, 1)}
    print(name, res[name], flush=True)
    del a, w, out
print(json.dumps(res))
