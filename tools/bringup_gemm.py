"""GPU bring-up for the tcgen05 GEMM: correctness vs torch.matmul over a shape sweep, then timing.

Run on a B200 (gpurun).  Each case runs in its own subprocess under a timeout so that a trap in one
case (poisoned context) cannot hide the others.
"""
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "advanced-video-event-detection-extraction_b200", "csrc", "libb200clip.so")


class Cfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("image_size", "patch", "width", "layers", "heads", "mlp_dim", "embed_dim", "act")] + \
               [("ln_eps", ctypes.c_float)] + \
               [(n, ctypes.c_int32) for n in
                ("text_ctx", "text_vocab", "text_width", "text_heads", "text_layers", "text_mlp_dim")]


def make_handle(lib):
    cfg = Cfg(224, 32, 768, 12, 12, 3072, 512, 0, 1e-5, 77, 49408, 512, 8, 12, 2048)
    h = ctypes.c_void_p()
    rc = lib.b200clip_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    if rc != 0:
        lib.b200clip_last_error.restype = ctypes.c_char_p
        raise RuntimeError(f"create failed {rc}: {lib.b200clip_last_error(None)}")
    return h


def run_case(m, n, k, bias, resid, act, timing):
    import torch
    lib = ctypes.CDLL(LIB)
    lib.b200clip_last_error.restype = ctypes.c_char_p
    h = make_handle(lib)
    torch.manual_seed(m * 7 + n * 3 + k)
    dev = "cuda:0"
    a = (torch.randn(m, k, device=dev) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
    b = torch.randn(n, device=dev, dtype=torch.float32) if bias else None
    r = torch.randn(m, n, device=dev).bfloat16() if resid else None
    out = torch.full((m, n), float("nan"), device=dev, dtype=torch.bfloat16)
    if resid:
        out.copy_(r)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        rc = lib.b200clip_gemm_bf16(h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(w.data_ptr()),
                                    ctypes.c_void_p(out.data_ptr()), m, n, k,
                                    ctypes.c_void_p(b.data_ptr() if bias else 0),
                                    ctypes.c_void_p(out.data_ptr() if resid else 0), act, ctypes.c_void_p(st))
        if rc != 0:
            raise RuntimeError(f"gemm rc={rc}: {lib.b200clip_last_error(h)}")

    call()
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
    got = out.float()
    err = (got - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()
    bad = (err > tol) | torch.isnan(got)
    res = {"m": m, "n": n, "k": k, "bias": bias, "resid": resid, "act": act,
           "max_err": float(err.nan_to_num(1e9).max()), "n_bad": int(bad.sum()), "n_nan": int(torch.isnan(got).sum())}
    if res["n_bad"]:
        # structure of the failure: which rows / columns are wrong
        rows_bad = bad.any(dim=1).nonzero().flatten()
        cols_bad = bad.any(dim=0).nonzero().flatten()
        res["rows_bad"] = [int(rows_bad.numel()), rows_bad[:12].tolist()]
        res["cols_bad"] = [int(cols_bad.numel()), cols_bad[:12].tolist()]
        res["sample"] = [[float(got[i, j]), float(ref[i, j])] for i, j in [(0, 0), (0, 1), (1, 0), (8, 0), (0, 8), (31, 40), (64, 100 % n), (127 % m, n - 1)]]
    if timing and not res["n_bad"]:
        if resid:
            out.copy_(r)
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
        res["ms"] = ms
        res["tflops"] = 2.0 * m * n * k / ms / 1e9
        # cuBLAS for context
        for _ in range(3):
            torch.matmul(a, w.t())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(a, w.t())
        e1.record()
        torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * m * n * k / (e0.elapsed_time(e1) / iters) / 1e9
    lib.b200clip_destroy(h)
    return res


CASES = [
    # m, n, k, bias, resid, act, timing
    (128, 256, 64, 0, 0, 0, 0),
    (128, 256, 256, 0, 0, 0, 0),
    (128, 128, 128, 0, 0, 0, 0),
    (256, 512, 768, 1, 0, 0, 0),
    (200, 768, 768, 1, 1, 0, 0),      # M tail
    (1000, 3072, 768, 1, 0, 1, 0),
    (1000, 768, 3072, 1, 1, 0, 0),
    (333, 2304, 768, 1, 0, 2, 0),
    (77, 512, 512, 1, 0, 0, 0),
    (77, 1536, 512, 1, 0, 0, 0),      # N=1536 -> BLOCK_N 256
    (130, 384, 512, 1, 0, 0, 0),      # N=384 -> BLOCK_N 128
    (51200, 2304, 768, 1, 0, 0, 1),
    (51200, 768, 768, 1, 1, 0, 1),
    (51200, 3072, 768, 1, 0, 1, 1),
    (51200, 768, 3072, 1, 1, 0, 1),
    (50176, 768, 3072, 0, 0, 0, 1),
    (8192, 8192, 8192, 0, 0, 0, 1),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        args = [int(x) for x in sys.argv[2:9]]
        print("RESULT " + json.dumps(run_case(*args)), flush=True)
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for c in CASES:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case"] + [str(x) for x in c], capture_output=True,
                               text=True, timeout=120)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                r = json.loads(line[0][7:])
            else:
                r = {"case": c, "rc": p.returncode, "stdout": p.stdout[-1500:], "stderr": p.stderr[-1500:]}
        except subprocess.TimeoutExpired:
            r = {"case": c, "timeout": True}
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
        results.append(r)
    with open(os.path.join(ROOT, "gpurun_out", "bringup_gemm.json"), "w") as f:
        json.dump(results, f, indent=1)
