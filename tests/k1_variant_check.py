"""Helper run in a SUBPROCESS by tests/test_gpu_preprocess.py::test_k1_code_path_variants: K1 picks its code path once
per process from environment switches (B200CLIP_AREA_FP32: fp32 area arithmetic instead of the integer-exact one;
B200CLIP_K1_UNFUSED: separate area / horizontal kernels; B200CLIP_AREA_NOSTRIP: per-pixel area kernel;
B200CLIP_AREA_HFIRST: horizontal-first integer area kernel instead of the vertical-first one;
B200CLIP_VPASS_GENERIC: per-item vertical-pass kernel instead of the tile form for the bf16 patch output;
B200CLIP_AREA_NOMMA: the CUDA-core kernels instead of the integer tensor-core (IMMA) form of stages A + B), so each
variant needs its own process.  Checks 1080p and 720p frames byte for byte against the oracle (which is itself pinned
to cv2 / Pillow / torchvision)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from b200clip import capi  # noqa: E402
from b200clip import open_clip as oc  # noqa: E402
from oracle import clip_ref  # noqa: E402
from oracle import preprocess_ref as P  # noqa: E402
from synth import noise_frames, structured_frames  # noqa: E402


def main() -> int:
    cfg = clip_ref.CONFIGS["ViT-B-32"]
    sd = clip_ref.init_state_dict(cfg, seed=0)
    model, _, _ = oc.create_model_and_transforms("ViT-B-32", state_dict=sd, device="cuda:0", max_images=8, max_texts=1)
    bad = 0
    for (w, h) in [(1920, 1080), (1280, 720), (1000, 700)]:
        frames = np.concatenate([noise_frames(2, h, w, seed=w + h), structured_frames(1, h, w, seed=w * 3 + h)])
        # extreme bytes exercise the rounding boundaries of the area stage
        frames[0, ::2] = 255
        frames[0, :, ::3] = 0
        chw = model.preprocess_u8(torch.from_numpy(frames).cuda(), capi.RESIZE_REFERENCE, chw=True).cpu().numpy()
        patches = model.preprocess_u8(torch.from_numpy(frames).cuda(), capi.RESIZE_REFERENCE, chw=False).float().cpu().numpy()
        for i in range(len(frames)):
            want = P.to_chw_normalized(P.reference_preprocess_u8(frames[i]))
            n = int((chw[i].view(np.uint32) != want.view(np.uint32)).sum())
            if n:
                print(f"{w}x{h} frame {i}: {n} of {want.size} values differ")
                bad += 1
            want_patches = torch.from_numpy(P.patchify(want, 32)).bfloat16().float().numpy()
            if not np.array_equal(patches[i * 49:(i + 1) * 49], want_patches):
                print(f"{w}x{h} frame {i}: bf16 patch rows differ")
                bad += 1
    # a batch large enough that every persistent CTA walks several (strip, frame) items with the ring running across them:
    # 600 frames of 720p (4 distinct) -> 7200 items on <= 444 resident CTAs; every copy must give the same bytes
    base = np.concatenate([noise_frames(3, 720, 1280, seed=5), structured_frames(1, 720, 1280, seed=6)])
    many = torch.from_numpy(base).cuda()[torch.arange(600, device="cuda") % 4]
    pm = model.preprocess_u8(many, capi.RESIZE_REFERENCE, chw=False).view(600, -1)
    first = model.preprocess_u8(torch.from_numpy(base).cuda(), capi.RESIZE_REFERENCE, chw=False).view(4, -1)
    if not torch.equal(pm.view(torch.int16), first[torch.arange(600, device="cuda") % 4].view(torch.int16)):
        print("600-frame batch: copies of a frame differ")
        bad += 1
    for i in range(4):
        want = torch.from_numpy(P.patchify(P.to_chw_normalized(P.reference_preprocess_u8(base[i])), 32)).bfloat16().reshape(-1)
        if not torch.equal(first[i].cpu().view(torch.int16), want.view(torch.int16)):
            print(f"720p frame {i}: bf16 patch rows differ from the oracle")
            bad += 1
    print("variant ok" if not bad else "variant FAILED", {k: v for k, v in os.environ.items() if k.startswith("B200CLIP_")})
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
