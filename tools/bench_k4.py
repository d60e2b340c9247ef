"""K4 (similarity + top-k + threshold + intervals) micro-bench over the BASELINE config sizes, one JSON line each:
achieved HBM GB/s (algorithmic bytes of SURVEY.md section 8d: n*E*sizeof + Q*E*4 + Q*k*12) and TFLOP/s, against
MEASURED_PEAKS.json.  Inputs are larger than the 126 MB L2 except where noted.  Results are checked against torch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from b200clip import capi
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config

peaks = {"hbm_gbs": 6555.2, "bf16_tflops_sustained": 1371.0}
try:
    peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except OSError:
    pass
h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
st = capi.c_void_p(torch.cuda.current_stream().cuda_stream)
CASES = [("config 1/2 scale: 3600 x 1 query fp32 (L2 resident)", 3600, 1, 5, torch.float32),
         ("config 5: 100k crops x 1 image query fp32, top-10", 100_000, 1, 10, torch.float32),
         ("1M x 1 query fp32", 1_000_000, 1, 5, torch.float32),
         ("1M x 1 query bf16", 1_000_000, 1, 5, torch.bfloat16),
         ("1M x 8 queries bf16", 1_000_000, 8, 5, torch.bfloat16),
         ("config 4: 1M x 256 queries bf16", 1_000_000, 256, 5, torch.bfloat16),
         ("config 4 (fp32 cache): 1M x 256 queries fp32", 1_000_000, 256, 5, torch.float32)]
e = 512
for name, n, q, k, dt in CASES:
    g = torch.Generator(device="cuda").manual_seed(n + q)
    img = torch.randn(n, e, device="cuda", generator=g)
    img = (img / img.norm(dim=-1, keepdim=True)).to(dt)
    txt = torch.randn(q, e, device="cuda", generator=g)
    txt = txt / txt.norm(dim=-1, keepdim=True)
    s = torch.empty(q, k, device="cuda")
    i = torch.empty(q, k, device="cuda", dtype=torch.int64)
    iv = torch.empty(q, k, 2, device="cuda", dtype=torch.float64)
    c = torch.empty(q, device="cuda", dtype=torch.int32)

    def call():
        h.call("b200clip_sim_topk", capi._p(img), capi.BF16 if dt == torch.bfloat16 else capi.F32, n, e, capi._p(txt), q, k,
               -1.0, None, 0, 30.0, 0.0, capi._p(s), capi._p(i), capi._p(iv), capi._p(c), st)

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
    qq = min(q, 8)
    ref = torch.topk(img.float() @ (txt[:qq].to(dt).float() if dt == torch.bfloat16 and q >= 16 else txt[:qq]).t(), k, dim=0)
    esz = 2 if dt == torch.bfloat16 else 4
    bytes_ = n * e * esz + q * e * 4 + q * k * 12
    print(json.dumps({"case": name, "ms": round(ms, 4), "hbm_gbs_algorithmic": round(bytes_ / ms / 1e6, 1),
                      "hbm_frac_of_measured_peak": round(bytes_ / ms / 1e6 / peaks["hbm_gbs"], 3),
                      "tflops": round(2.0 * n * q * e / ms / 1e9, 1),
                      "top1_matches_torch": float((i[:qq, 0] == ref.indices[0]).float().mean())}), flush=True)
    del img
