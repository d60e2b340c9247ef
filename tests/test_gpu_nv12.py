"""Frame feed: NV12 decoder output converted inside K1.  The bar is byte-exactness: the patch
rows / CHW tensors K1 produces from an NV12 frame must equal what the RGB chain produces from
cv2.cvtColor(frame, COLOR_YUV2RGB_NV12) of the same frame (the installed OpenCV is the pin; oracle/nv12_ref.py restates
it and is itself pinned in tests/test_oracle_nv12.py), and through it the oracle's reference preprocess."""
import os

import numpy as np
import pytest
import torch

from synth import structured_frames

pytestmark = pytest.mark.gpu


def _nv12_batch(n, h, w, seed):
    """Half noise (every tap and every saturation branch matters), half structured content through a BT.601 encode."""
    from oracle import nv12_ref

    rng = np.random.default_rng(seed)
    out = rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    for i, f in enumerate(structured_frames(n // 2, h, w, seed=seed + 1)):
        out[2 * i + 1] = nv12_ref.rgb_to_nv12(f)
    return out


def _rgb(nv):
    import cv2

    return np.stack([cv2.cvtColor(f, cv2.COLOR_YUV2RGB_NV12) for f in nv])


# fused kernel: 1080p, 720p, 960x540 (integer-exact area geometries, 16-byte aligned planes); everything else converts the
# crop window first: non-integer-exact scale, width not a multiple of 16, no area stage at all, already 224x224
GEOMETRIES = [(1080, 1920), (720, 1280), (540, 960), (700, 1000), (562, 1000), (480, 640), (360, 492), (224, 224), (226, 300)]


@pytest.mark.parametrize("h,w", GEOMETRIES)
def test_patches_equal_cv2_then_rgb_chain(model_b32, h, w):
    from b200clip import capi
    from oracle import nv12_ref
    from oracle import preprocess_ref as P

    nv = _nv12_batch(4, h, w, seed=h + w)
    rgb = _rgb(nv)
    assert np.array_equal(rgb[0], nv12_ref.nv12_to_rgb(nv[0]))
    dev = torch.from_numpy(nv).cuda()
    for mode in (capi.RESIZE_REFERENCE, capi.RESIZE_BICUBIC, capi.RESIZE_BILINEAR_AA):
        got = model_b32.preprocess_nv12(dev, mode)
        want = model_b32.preprocess_u8(torch.from_numpy(rgb).cuda(), mode)
        assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (h, w, mode)
    chw = model_b32.preprocess_nv12(dev, capi.RESIZE_REFERENCE, chw=True).cpu().numpy()
    for i in (0, 1):
        want = P.to_chw_normalized(P.reference_preprocess_u8(rgb[i]))
        assert np.array_equal(chw[i].view(np.uint32), want.view(np.uint32)), (h, w, i)


def test_saturation_extremes_1080p(model_b32):
    """Y / U / V combinations that drive every channel below 0 and above 255, in blocks large enough to survive the
    area shrink, plus a checkerboard of extremes."""
    from b200clip import capi

    h, w = 1080, 1920
    nv = np.empty((2, h * 3 // 2, w), np.uint8)
    vals = np.array([0, 16, 128, 235, 255], np.uint8)
    yy = vals[(np.arange(h)[:, None] // 40 + np.arange(w)[None, :] // 56) % 5]
    uu = vals[(np.arange(h // 2)[:, None] // 12 + np.arange(w // 2)[None, :] // 20) % 5]
    vv = vals[(np.arange(h // 2)[:, None] // 28 + 2 * (np.arange(w // 2)[None, :] // 36)) % 5]
    nv[0, :h] = yy
    nv[0, h:, 0::2], nv[0, h:, 1::2] = uu, vv
    nv[1] = np.where((np.indices((h * 3 // 2, w)).sum(0) & 1) == 0, 0, 255)
    got = model_b32.preprocess_nv12(torch.from_numpy(nv).cuda(), capi.RESIZE_REFERENCE)
    want = model_b32.preprocess_u8(torch.from_numpy(_rgb(nv)).cuda(), capi.RESIZE_REFERENCE)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))


def test_pitched_planes_like_a_decoder(model_b32):
    """NVDEC-style surface: row pitch > width, chroma plane behind an aligned luma height, per-frame strides."""
    from b200clip import capi

    h, w, pitch, hal = 720, 1280, 1536, 736
    nv = _nv12_batch(3, h, w, seed=77)
    surf = torch.zeros(3, hal * 3 // 2, pitch, dtype=torch.uint8)
    surf[:, :h, :w] = torch.from_numpy(nv[:, :h])
    surf[:, hal:hal + h // 2, :w] = torch.from_numpy(nv[:, h:])
    dev = surf.cuda()
    g = model_b32.cfg.image_size // model_b32.cfg.patch
    out = torch.empty(3 * g * g, model_b32.cfg.patch_k, device="cuda", dtype=torch.bfloat16)
    fs = hal * 3 // 2 * pitch
    model_b32.handle.call("b200clip_preprocess_nv12", capi._p(dev.data_ptr()), capi._p(dev.data_ptr() + hal * pitch), 3, h, w,
                          fs, fs, pitch, capi.RESIZE_REFERENCE, capi._p(out), capi._p(None), model_b32._stream())
    want = model_b32.preprocess_u8(torch.from_numpy(_rgb(nv)).cuda(), capi.RESIZE_REFERENCE)
    assert torch.equal(out.view(torch.int16), want.view(torch.int16))
    # an unaligned plane pointer must take the generic path and still be exact
    shifted = torch.zeros(fs * 3 + 64, dtype=torch.uint8, device="cuda")
    shifted[4:4 + fs * 3] = dev.reshape(-1)
    out2 = torch.empty_like(out)
    model_b32.handle.call("b200clip_preprocess_nv12", capi._p(shifted.data_ptr() + 4), capi._p(shifted.data_ptr() + 4 + hal * pitch),
                          3, h, w, fs, fs, pitch, capi.RESIZE_REFERENCE, capi._p(out2), capi._p(None), model_b32._stream())
    assert torch.equal(out2.view(torch.int16), want.view(torch.int16))


def test_bad_arguments_are_rejected(model_b32):
    from b200clip import capi

    d = torch.zeros(1, 1080 * 3 // 2, 1920, dtype=torch.uint8, device="cuda")
    out = torch.empty(49, model_b32.cfg.patch_k, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(capi.B200ClipError):      # odd height
        model_b32.handle.call("b200clip_preprocess_nv12", capi._p(d), capi._p(d), 1, 1079, 1920, 1920 * 1620, 1920 * 1620, 1920,
                              0, capi._p(out), capi._p(None), model_b32._stream())
    with pytest.raises(capi.B200ClipError):      # pitch smaller than the width
        model_b32.handle.call("b200clip_preprocess_nv12", capi._p(d), capi._p(d), 1, 1080, 1920, 1920 * 1620, 1920 * 1620, 1000,
                              0, capi._p(out), capi._p(None), model_b32._stream())
    with pytest.raises(capi.B200ClipError):      # no output
        model_b32.handle.call("b200clip_preprocess_nv12", capi._p(d), capi._p(d), 1, 1080, 1920, 1920 * 1620, 1920 * 1620, 1920,
                              0, capi._p(None), capi._p(None), model_b32._stream())
    with pytest.raises(ValueError):
        model_b32.preprocess_nv12(torch.zeros(1, 1081, 1920, dtype=torch.uint8, device="cuda"))
    model_b32.handle.call("b200clip_encode_frames_nv12", capi._p(None), capi._p(None), 0, 1080, 1920, 0, 0, 0, 0, capi._p(None),
                          capi.F32, 1, model_b32._stream())      # n = 0 is a no-op


@pytest.mark.parametrize("h,w,n", [(1080, 1920, 37), (720, 1280, 20), (562, 1000, 9)])
def test_embeddings_device_and_host_paths(model_b32, h, w, n):
    """encode_frames_nv12 (device frames) and encode_frames_nv12_host (pinned and pageable host frames, windowed
    upload) give bit-identical embeddings to the RGB entry points on the cv2-converted frames, at half the H2D bytes."""
    from b200clip import capi

    nv = _nv12_batch(n, h, w, seed=n)
    rgb = _rgb(nv)
    want = model_b32.encode_frames_u8(torch.from_numpy(rgb).cuda(), capi.RESIZE_REFERENCE, normalize=True)
    got = model_b32.encode_frames_nv12(torch.from_numpy(nv).cuda(), capi.RESIZE_REFERENCE, normalize=True)
    assert torch.equal(got, want)
    model_b32.handle.transfer_bytes(reset=True)
    host = model_b32.encode_frames_nv12_host(nv, capi.RESIZE_REFERENCE, True)
    h2d_nv, _ = model_b32.handle.transfer_bytes(reset=True)
    assert np.array_equal(host, want.cpu().numpy())
    pinned = torch.from_numpy(nv).pin_memory()
    dev_out = torch.empty(n, model_b32.embed_dim, device="cuda")
    model_b32.encode_frames_nv12_host(pinned, capi.RESIZE_REFERENCE, True, out=dev_out)
    torch.cuda.synchronize()
    assert torch.equal(dev_out, want)
    model_b32.handle.transfer_bytes(reset=True)
    host_rgb = model_b32.encode_frames_u8_host(rgb, capi.RESIZE_REFERENCE, True)
    h2d_rgb, _ = model_b32.handle.transfer_bytes(reset=True)
    assert np.array_equal(host_rgb, want.cpu().numpy())
    assert h2d_nv <= 0.52 * h2d_rgb and h2d_nv < n * h * w * 1.5, (h2d_nv, h2d_rgb)
    bf = model_b32.encode_frames_nv12(torch.from_numpy(nv).cuda(), capi.RESIZE_REFERENCE, True, torch.bfloat16)
    assert torch.equal(bf, want.bfloat16())


def test_unfused_variant_in_subprocess():
    """B200CLIP_NV12_UNFUSED=1 forces the generic window conversion for the fused geometries too: both forms must agree
    with cv2 (the form is chosen once per process)."""
    import subprocess
    import sys

    code = r'''
import sys, numpy as np, torch, cv2
sys.path.insert(0, %r); sys.path.insert(0, %r)
from b200clip import capi
from b200clip import open_clip as oc
from oracle import clip_ref
cfg = clip_ref.CONFIGS["ViT-B-32"]
model, _, _ = oc.create_model_and_transforms("ViT-B-32", state_dict=clip_ref.init_state_dict(cfg, seed=0), device="cuda:0", max_images=8)
for h, w in [(1080, 1920), (720, 1280)]:
    nv = np.random.default_rng(h).integers(0, 256, (3, h * 3 // 2, w), dtype=np.uint8)
    rgb = np.stack([cv2.cvtColor(f, cv2.COLOR_YUV2RGB_NV12) for f in nv])
    got = model.preprocess_nv12(torch.from_numpy(nv).cuda(), capi.RESIZE_REFERENCE)
    want = model.preprocess_u8(torch.from_numpy(rgb).cuda(), capi.RESIZE_REFERENCE)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (h, w)
print("variant ok", model.handle.launches)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    launches = {}
    for name, env in (("fused", {}), ("unfused", {"B200CLIP_NV12_UNFUSED": "1"}), ("persistent", {"B200CLIP_K1_PERSISTENT": "1"})):
        e = {k: v for k, v in os.environ.items() if k not in ("B200CLIP_NV12_UNFUSED", "B200CLIP_K1_PERSISTENT")}
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "variant ok" in r.stdout, (name, r.stdout[-1500:], r.stderr[-1500:])
        launches[name] = int(r.stdout.strip().split()[-1])
    assert launches["unfused"] > launches["fused"]       # the generic form adds a conversion launch per call
