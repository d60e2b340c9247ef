"""2-CTA GEMM with one pair per cluster (B200CLIP_GEMM_PAIRS unset) vs two pairs sharing their B tile by TMA multicast
(B200CLIP_GEMM_PAIRS=2): correctness against an fp32 torch reference on the first / last rows, and time per shape.
usage: [B200CLIP_GEMM_PAIRS=2] python tools/probes/gemm_pairs_probe.py [M]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from b200clip import capi
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config

h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 180000
shapes = [("qkv", M, 2304, 768, 1, 0, 0), ("out", M, 768, 768, 1, 1, 0), ("fc", M, 3072, 768, 1, 0, 1),
          ("proj", M, 768, 3072, 1, 1, 0), ("tail", 8192 + 300, 768, 768, 1, 0, 0)]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {"pairs": os.environ.get("B200CLIP_GEMM_PAIRS", "1")}
torch.manual_seed(0)
for name, m, n, k, bias, resid, act in shapes:
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    b = torch.randn(n, device="cuda") if bias else None
    r = (torch.randn(m, n, device="cuda") * 0.5).bfloat16() if resid else None
    out = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)

    def call(o):
        h.call("b200clip_gemm_bf16", capi._p(a), capi._p(w), capi._p(o), m, n, k, capi._p(b), capi._p(r), act, st)

    call(out)
    torch.cuda.synchronize()
    err = 0.0
    for lo, hi in ((0, min(m, 1536)), (max(0, m - 1536), m), (m // 2, min(m, m // 2 + 600))):
        ref = a[lo:hi].float() @ w.float().t()
        if bias:
            ref += b
        if act == 1:
            ref = ref * torch.sigmoid(1.702 * ref)
        if resid:
            ref += r[lo:hi].float()
        err = max(err, float((out[lo:hi].float() - ref).abs().max() / (ref.abs().max() + 1e-6)))
    tmp = torch.empty_like(out)
    for _ in range(3):
        call(tmp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call(tmp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[name] = {"ms": round(ms, 4), "tflops": round(2.0 * m * n * k / ms / 1e9, 1), "rel_err": round(err, 5)}
    print(name, res[name], flush=True)
    del a, w, out, tmp
print(json.dumps(res))
