"""GPU parity of the building-block kernels through the C ABI (b200clip_gemm_bf16 / layernorm / attention)
against plain PyTorch fp32 references of the same op on the same seeded inputs."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from b200clip import capi
    from b200clip.model_configs import MODEL_CONFIGS, to_capi_config

    h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
    yield h
    h.close()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("m,n,k,bias,resid,act", [
    (128, 256, 64, 0, 0, 0), (200, 768, 768, 1, 1, 0), (1000, 3072, 768, 1, 0, 1), (1000, 768, 3072, 1, 1, 0),
    (333, 2304, 768, 1, 0, 2), (77, 512, 512, 1, 0, 0), (130, 384, 512, 1, 0, 0), (50, 64, 64, 1, 0, 0),
    (5000, 768, 3072, 0, 0, 0), (1, 768, 768, 1, 0, 0), (12544, 1024, 640, 0, 0, 0),
    # M >= 2048 and N % 256 == 0 -> the 2-CTA (cta_group::2) kernel; tails in M, every epilogue
    (4096, 2304, 768, 1, 0, 0), (2049, 768, 768, 1, 1, 0), (3000, 3072, 768, 1, 0, 1), (2304, 256, 64, 0, 0, 2),
    (20000, 768, 3072, 1, 1, 0), (2048, 512, 2048, 1, 1, 0),
])
def test_gemm_matches_torch(handle, m, n, k, bias, resid, act):
    from b200clip import capi

    torch.manual_seed(m + n + k)
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    b = torch.randn(n, device="cuda") if bias else None
    r = torch.randn(m, n, device="cuda").bfloat16() if resid else None
    out = torch.full((m, n), float("nan"), device="cuda", dtype=torch.bfloat16)
    if resid:
        out.copy_(r)
    handle.call("b200clip_gemm_bf16", capi._p(a), capi._p(w), capi._p(out), m, n, k, capi._p(b),
                capi._p(out if resid else None), act, _stream())
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + r.float()
    err = (out.float() - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()   # bf16 output rounding (2^-9 relative) + accumulation order
    assert not torch.isnan(out.float()).any()
    assert bool((err <= tol).all()), f"max err {float(err.max())}"


@pytest.mark.parametrize("rows,width", [(1000, 768), (77, 512), (257, 1024), (5, 128), (9, 64)])
def test_layernorm_matches_torch(handle, rows, width):
    from b200clip import capi

    torch.manual_seed(rows)
    x = (torch.randn(rows, width, device="cuda") * 3 + 0.5).bfloat16()
    g = torch.randn(width, device="cuda") * 0.2 + 1
    b = torch.randn(width, device="cuda") * 0.1
    y = torch.empty_like(x)
    handle.call("b200clip_layernorm_bf16", capi._p(x), capi._p(g), capi._p(b), capi._p(y), rows, width, 1e-5, _stream())
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (width,), g, b, 1e-5)
    assert float((y.float() - ref).abs().max()) <= 0.02 + 0.008 * float(ref.abs().max())


@pytest.mark.parametrize("n_seq,t,heads,causal", [(3, 50, 12, 0), (2, 77, 8, 1), (2, 257, 16, 0), (5, 5, 2, 0),
                                                    (1, 16, 1, 1), (70, 50, 12, 0), (1, 64, 2, 1), (1, 65, 2, 0),
                                                    # >= 296 (sequence, head) items, T <= 64, no mask: persistent tcgen05 kernel, two
                                                    # items per 128-row tile (odd item counts: the last tile holds one)
                                                    (300, 50, 12, 0), (40, 64, 8, 0), (100, 33, 4, 0), (500, 1, 1, 0),
                                                    (37, 7, 9, 0), (301, 50, 1, 0), (2000, 50, 12, 0),
                                                    # 64 < T <= 320, no mask: K/V-resident kernel (one CTA per sequence and head)
                                                    (3, 257, 16, 0), (5, 100, 4, 0), (2, 320, 2, 0), (1, 129, 1, 0), (4, 96, 3, 0),
                                                    # beyond it / causal: the tiled kernel
                                                    (1, 321, 2, 0), (2, 200, 4, 1)])
def test_attention_matches_torch(handle, n_seq, t, heads, causal):
    from b200clip import capi

    torch.manual_seed(t * heads)
    d = heads * 64
    qkv = (torch.randn(n_seq * t, 3 * d, device="cuda") * 1.5).bfloat16()
    out = torch.full((n_seq * t, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    handle.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, causal, _stream())
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_seq, t, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((t, t), float("-inf"), device="cuda").triu_(1)
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(n_seq * t, d)
    assert not torch.isnan(out.float()).any()
    # P is rounded to bf16 before the PV product and the output is bf16
    assert float((out.float() - ref).abs().max()) <= 0.03


def test_attention_kernel_variants_in_subprocess():
    """The attention kernel for 64 < T <= 320 is chosen once per process: default = the persistent tcgen05 kernel with the
    probabilities in tensor memory for T = 257 (ViT-L/14) and mma.sync with resident K/V otherwise,
    B200CLIP_ATTN_NOTC2=1 = mma.sync also for T = 257, B200CLIP_ATTN_TC=1 = the first tcgen05 kernel (S and PV on the
    5th-generation tensor cores, V as an MN-major operand, P through shared memory), B200CLIP_ATTN_TILED=1 = the tiled
    fallback; for T <= 64 the default is the two-items-per-tile tcgen05 kernel, B200CLIP_ATTN_NOTC64=1 = the persistent
    mma.sync kernel, B200CLIP_ATTN_ONESHOT=1 = one CTA per item.  All must match torch on the ViT-L/14 and ViT-B/32
    shapes (few and many items per persistent CTA, odd item counts) and friends."""
    import os
    import subprocess
    import sys

    code = r'''
import ctypes, sys, torch
sys.path.insert(0, %r)
from b200clip import capi
from b200clip.model_configs import MODEL_CONFIGS, to_capi_config
h = capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for n_seq, t, heads in [(3, 257, 16), (1, 257, 1), (37, 257, 16), (300, 50, 12), (301, 50, 1), (77, 64, 5), (150, 33, 3), (5, 100, 4), (2, 320, 2), (1, 129, 1), (4, 96, 3), (2, 65, 2), (7, 145, 5), (2, 130, 2),
                          (2, 196, 3), (1, 258, 1), (3, 197, 12), (2, 260, 2), (1, 68, 1)]:
    torch.manual_seed(t * heads)
    d = heads * 64
    qkv = (torch.randn(n_seq * t, 3 * d, device="cuda") * 1.5).bfloat16()
    out = torch.full((n_seq * t, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    h.call("b200clip_attention_bf16", capi._p(qkv), capi._p(out), n_seq, t, heads, 0, st)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_seq, t, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(n_seq * t, d)
    err = float((out.float() - ref).abs().max())
    assert not torch.isnan(out.float()).any() and err <= 0.03, (n_seq, t, heads, err)
print("variant ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env in ({}, {"B200CLIP_ATTN_NOTC2": "1", "B200CLIP_ATTN_NOTC64": "1"}, {"B200CLIP_ATTN_TC": "1", "B200CLIP_ATTN_NOTC64": "1", "B200CLIP_ATTN_ONESHOT": "1"},
                {"B200CLIP_ATTN_TILED": "1"}):
        e = {k: v for k, v in os.environ.items() if not k.startswith("B200CLIP_ATTN")}
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "variant ok" in r.stdout, (env, r.stdout[-1500:], r.stderr[-1500:])


def test_head_kernel_variants_agree_in_subprocess():
    """ln_post -> proj -> L2 (SURVEY a4/a5 tail) runs on head_mma_kernel (mma.sync, LayerNorm output split into two bf16
    planes) by default and on the fp32 CUDA-core head_kernel with B200CLIP_HEAD_SIMT=1; the switch is read once per
    process, so tools/head_check.py embeds the same frames / texts (1, 5, 33, 70 rows: partial and several CTAs; both
    towers; normalised and raw; ViT-B/32 and ViT-L/14 widths) in one process per form and compares to 2e-5."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = {k: v for k, v in os.environ.items() if k != "B200CLIP_HEAD_SIMT"}
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "head_check.py")], env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])


def test_gemm_five_stage_form_in_subprocess():
    """GEMMs without a residual run the six-stage / one-staging-box form of the 2-CTA kernel by default;
    B200CLIP_GEMM_5STAGE=1 (read once per process) sends them through the five-stage / two-box form that residual GEMMs
    use.  Same parametrised comparison against torch, in its own process."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, B200CLIP_GEMM_5STAGE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_kernels.py"), "-m", "gpu", "-x", "-q",
                        "-k", "test_gemm_matches_torch"], env=e, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
