"""Development check of head_mma_kernel against the CUDA-core head kernel: embeddings of the same frames / texts from
two processes (B200CLIP_HEAD_SIMT=1 and default) must agree to ~1e-6; prints per-batch maxima of |difference|."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from b200clip import capi
    from b200clip import open_clip as oc
    from b200clip.model_configs import MODEL_CONFIGS
    from b200clip.weights import random_state_dict
    outs = {}
    for name in ("ViT-B-32", "ViT-L-14"):
        cfg = MODEL_CONFIGS[name]
        model, _, _ = oc.create_model_and_transforms(name, state_dict=random_state_dict(cfg, 0), device="cuda:0", max_images=64, max_texts=8)
        g = torch.Generator(device="cuda").manual_seed(3)
        for n in (1, 5, 33, 70):
            fr = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g)
            outs[f"{name}_img{n}"] = model.encode_frames_u8(fr, capi.RESIZE_BICUBIC, normalize=True).float().cpu().numpy()
            outs[f"{name}_raw{n}"] = model.encode_frames_u8(fr, capi.RESIZE_BICUBIC, normalize=False).float().cpu().numpy()
        tok = torch.randint(1, 1000, (5, 77), device="cuda", generator=g)
        tok[:, 0] = 49406; tok[:, 20] = 49407
        outs[f"{name}_txt"] = model.encode_text(tok, normalize=True).float().cpu().numpy()
    np.savez(sys.argv[2], **outs)
    sys.exit(0)
import numpy as np
for tag, env in (("simt", {"B200CLIP_HEAD_SIMT": "1"}), ("mma", {})):
    e = dict(os.environ); e.update(env)
    subprocess.run([sys.executable, __file__, "worker", f"/tmp/head_{tag}.npz"], env=e, check=True)
a, b = np.load("/tmp/head_simt.npz"), np.load("/tmp/head_mma.npz")
bad = 0
for k in a.files:
    d = float(np.abs(a[k] - b[k]).max()); s = float(np.abs(a[k]).max())
    print(k, "max |diff|", d, "max |value|", round(s, 4), "nan", int(np.isnan(b[k]).sum()))
    bad += d > 2e-5 * max(s, 1.0) or np.isnan(b[k]).any()
sys.exit(1 if bad else 0)
