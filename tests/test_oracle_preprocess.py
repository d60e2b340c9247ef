"""Pins oracle/preprocess_ref.py: byte-for-byte against the installed cv2 / Pillow / torchvision (the libraries the
reference itself calls) and against the golden fixtures produced through the reference's own call chain."""
import os

import numpy as np
import pytest

from oracle import preprocess_ref as P
from synth import noise_frames, structured_frames

cv2 = pytest.importorskip("cv2")
PIL_Image = pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 512, 288), (1280, 720, 512, 288), (800, 600, 512, 384),
                                         (2048, 1024, 512, 256), (1024, 1024, 512, 512), (640, 480, 512, 384),
                                         (1536, 768, 512, 256)])
def test_inter_area_matches_cv2(sw, sh, dw, dh):
    img = noise_frames(1, sh, sw, seed=sw)[0]
    assert P.fit_size(sw, sh) == (dw, dh)
    assert np.array_equal(P.inter_area_resize(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("w,h", [(512, 288), (1920, 1080), (300, 300), (224, 224), (288, 512), (399, 224), (225, 400)])
def test_clip_transform_matches_torchvision(w, h):
    import torchvision.transforms as T
    from torchvision.transforms import InterpolationMode

    img = noise_frames(1, h, w, seed=w * 7 + h)[0]
    tf = T.Compose([T.Resize(224, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(224), T.ToTensor(),
                    T.Normalize(P.OPENAI_MEAN, P.OPENAI_STD)])
    want = tf(PIL_Image.fromarray(img)).numpy()
    got = P.to_chw_normalized(P.clip_transform_u8(img))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_golden_chain(golden_dir):
    g = np.load(os.path.join(golden_dir, "vitb32_1080p.npz"))
    hd = np.concatenate([structured_frames(4, 1080, 1920, seed=7), noise_frames(2, 1080, 1920, seed=8)])
    for i in (0, 4):
        assert np.array_equal(P.reference_preprocess_u8(hd[i]), g["pre_u8"][i])
    geo = np.load(os.path.join(golden_dir, "preprocess_geometries.npz"))
    for key in geo.files:
        dims, seed = key.split("_s")
        w, h = (int(v) for v in dims.split("x"))
        assert np.array_equal(P.reference_preprocess_u8(noise_frames(1, h, w, seed=int(seed))[0]), geo[key]), key


def test_patchify_is_conv_unfold():
    import torch

    chw = np.random.default_rng(0).standard_normal((3, 64, 64)).astype(np.float32)
    w = torch.randn(16, 3, 32, 32)
    conv = torch.nn.functional.conv2d(torch.from_numpy(chw)[None], w, stride=32)[0].reshape(16, -1).T
    gemm = torch.from_numpy(P.patchify(chw, 32)) @ w.reshape(16, -1).T
    assert torch.allclose(conv, gemm, atol=1e-4)


def test_oracle_pipeline_matches_second_1080p_golden(golden_dir):
    """The CPU restatement (cv2 INTER_AREA shrink -> PIL transform -> fp32 ViT) against what the reference's own classes
    produced for the 24-frame 1080p golden (tests/golden/make_golden_1080p_more.py): same embeddings to fp32 noise, on
    a structured and a noise frame (a 1080p frame costs about a second on the CPU)."""
    import sys

    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    from make_golden_1080p_more import N_STRUCTURED, SEED_NOISE, SEED_STRUCTURED
    from oracle import clip_ref
    from oracle.reference_pipeline import ReferenceCPU
    from synth import noise_frames, structured_frames

    g = np.load(os.path.join(golden_dir, "vitb32_1080p_more.npz"))
    ref = ReferenceCPU("ViT-B-32", state_dict=clip_ref.init_state_dict(clip_ref.CONFIGS["ViT-B-32"], seed=0, gain=1.0))
    frames = np.concatenate([structured_frames(1, 1080, 1920, seed=SEED_STRUCTURED), noise_frames(1, 1080, 1920, seed=SEED_NOISE)])
    emb = ref.encode_images(frames, shrink=True)
    assert np.abs(emb[0] - g["emb"][0]).max() < 2e-5
    assert np.abs(emb[1] - g["emb"][N_STRUCTURED]).max() < 2e-5
