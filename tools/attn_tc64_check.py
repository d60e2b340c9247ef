"""Development check of attention_tc64_kernel (T <= 64, two items per tcgen05 tile; B200CLIP_ATTN_TC64=1 while opt-in)
against torch on several shapes (odd item counts, T < 64 masks, many tiles per CTA), then the per-call time on the
ViT-B/32 shape (3600 sequences x 12 heads, T = 50)."""
import ctypes, os, sys
os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from b200clip import capi
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config
h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
bad = 0
for n_seq, t, heads in [(200, 50, 12), (301, 50, 1), (77, 64, 5), (150, 33, 3), (400, 1, 1), (3600, 50, 12), (333, 17, 7)]:
    torch.manual_seed(n_seq * heads + t)
    d = heads * 64
    qkv = (torch.randn(n_seq * t, 3 * d, device="cuda") * 1.5).bfloat16()
    out = torch.full((n_seq * t, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    h.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, 0, st)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_seq, t, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(n_seq * t, d)
    nan = int(torch.isnan(out.float()).sum())
    err = float((out.float() - ref).nan_to_num(1e9).abs().max())
    print(n_seq, t, heads, "nan", nan, "max err", err, flush=True)
    bad += (nan > 0) or (err > 0.03)
n_seq, t, heads = 3600, 50, 12
qkv = (torch.randn(n_seq * t, 3 * heads * 64, device="cuda") * 1.5).bfloat16()
out = torch.empty(n_seq * t, heads * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    h.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, 0, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(48):
    h.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, 0, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 48
print("ms per layer call (3600 seq x 12 heads, T = 50):", round(ms, 4), "GB/s", round(n_seq * t * heads * 64 * 2 * 4 / ms / 1e6, 1),
      {k: v for k, v in os.environ.items() if k.startswith("B200CLIP_ATTN")})
sys.exit(1 if bad else 0)
