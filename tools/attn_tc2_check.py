"""Development check of attention_tc2_kernel (T = 257, the default; B200CLIP_ATTN_NOTC2=1 = mma.sync kernel) against torch, several batch sizes (fewer items than SMs, a
non-multiple, many items per CTA), then the per-call time of both kernels on the ViT-L/14 shape."""
import ctypes, os, sys
os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from b200clip import capi
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config
h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-L-14"]), 0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
bad = 0
for n_seq, heads in [(1, 1), (3, 16), (7, 5), (40, 16), (64, 16)]:
    t = 257
    torch.manual_seed(n_seq * heads)
    d = heads * 64
    qkv = (torch.randn(n_seq * t, 3 * d, device="cuda") * 1.5).bfloat16()
    out = torch.full((n_seq * t, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    h.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, 0, st)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_seq, t, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(n_seq * t, d)
    nan = int(torch.isnan(out.float()).sum())
    err = float((out.float() - ref).nan_to_num(1e9).abs().max())
    rows = (out.float() - ref).nan_to_num(1e9).abs().amax(dim=1).view(n_seq, t)
    print(n_seq, heads, "nan", nan, "max err", err, "worst rows", rows.amax(0).topk(3).indices.tolist(), flush=True)
    bad += (nan > 0) or (err > 0.03)
n_seq, t, heads = 512, 257, 16
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
print("ms per layer call (512 seq x 16 heads, T = 257):", round(e0.elapsed_time(e1) / 48, 4), {k: v for k, v in os.environ.items() if k.startswith("B200CLIP_ATTN")})
lib = capi.load_library()
if hasattr(lib, "b200clip_debug_a2_probe"):       # PROBES build: timeline of CTA 0, items 8 .. 15 (clock64, relative)
    buf = (ctypes.c_longlong * 192)()
    lib.b200clip_debug_a2_probe(buf, 192)
    v = [buf[i] for i in range(192)]
    t0 = min(x for x in v if x > 0)
    names = ["top", "full", "s'", "s_full", "pass1", "P", "o_full", "stored"]
    for j in range(8):
        for who, nm in ((0, "WG0"), (1, "WG1")):
            print(nm, "item", 8 + j, " ".join(f"{names[k]}={v[(who * 8 + j) * 8 + k] - t0}" for k in range(8)))
        print("MMA item", 8 + j, " ".join(f"{n}={v[(2 * 8 + j) * 8 + k] - t0}" for k, n in ((0, "S0"), (2, "S1"), (4, "PV0"), (6, "PV1"))))
sys.exit(1 if bad else 0)
