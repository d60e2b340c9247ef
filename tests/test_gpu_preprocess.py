"""K1 parity (byte-exact): the CUDA preprocess kernel against the numpy oracle (oracle/preprocess_ref.py) on seeded
frames, and against tests/golden/*.npz produced by the real cv2 / Pillow / torchvision through the reference's own
MemoryManager.resize_frame_for_memory + open_clip transform."""
import os

import numpy as np
import pytest
import torch

from synth import noise_frames, structured_frames

pytestmark = pytest.mark.gpu


def _u8_from_chw(chw: np.ndarray) -> np.ndarray:
    """invert the ToTensor/Normalize lookup: nearest table entry per channel -> uint8 HWC."""
    from oracle.preprocess_ref import normalize_table

    tab = normalize_table()
    out = np.stack([np.abs(tab[c][None, None, :] - chw[c][..., None]).argmin(-1) for c in range(3)], -1)
    return out.astype(np.uint8)


@pytest.mark.parametrize("w,h", [(224, 224), (512, 288), (1920, 1080), (1280, 720), (640, 480), (300, 300),
                                 (288, 512), (399, 224), (1024, 1024), (800, 600), (2048, 1024), (225, 400),
                                 (960, 540), (1600, 900)])   # integer-exact geometries with Dx = Dy = 15 / 25
def test_reference_mode_bit_exact_vs_oracle(model_b32, w, h):
    from b200clip import capi
    from oracle import preprocess_ref as P

    frames = np.concatenate([noise_frames(1, h, w, seed=w + h), structured_frames(1, h, w, seed=w * 3 + h)])
    dev = torch.from_numpy(frames).cuda()
    chw = model_b32.preprocess_u8(dev, capi.RESIZE_REFERENCE, chw=True).cpu().numpy()
    patches = model_b32.preprocess_u8(dev, capi.RESIZE_REFERENCE, chw=False).float().cpu().numpy()
    for i in range(len(frames)):
        want_u8 = P.reference_preprocess_u8(frames[i])
        want = P.to_chw_normalized(want_u8)
        assert np.array_equal(chw[i].view(np.uint32), want.view(np.uint32)), \
            f"{w}x{h} frame {i}: {(chw[i] != want).sum()} of {want.size} values differ"
        want_patches = torch.from_numpy(P.patchify(want, 32)).bfloat16().float().numpy()
        assert np.array_equal(patches[i * 49:(i + 1) * 49], want_patches)


@pytest.mark.parametrize("w,h", [(512, 288), (640, 480), (1000, 700)])
def test_transform_only_mode_matches_pil(model_b32, w, h):
    """RESIZE_BICUBIC == open_clip's transform alone (no 512 shrink), checked against the real torchvision/Pillow."""
    from PIL import Image

    from b200clip import capi
    from oracle.open_clip_shim import image_transform

    frame = noise_frames(1, h, w, seed=5)[0]
    want = image_transform(224)(Image.fromarray(frame)).numpy()
    got = model_b32.preprocess_u8(torch.from_numpy(frame[None]).cuda(), capi.RESIZE_BICUBIC, chw=True)[0].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_bilinear_mode_matches_pil_bilinear(model_b32):
    from PIL import Image

    from b200clip import capi
    from oracle import preprocess_ref as P

    frame = noise_frames(1, 1080, 1920, seed=6)[0]
    nw, nh = P.resize_output_size(1920, 1080)
    r = np.asarray(Image.fromarray(frame).resize((nw, nh), Image.BILINEAR))
    left, top = P.center_crop_box(nw, nh)
    want = P.to_chw_normalized(r[top:top + 224, left:left + 224])
    got = model_b32.preprocess_u8(torch.from_numpy(frame[None]).cuda(), capi.RESIZE_BILINEAR_AA, chw=True)[0].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_golden_1080p_and_geometries(model_b32, golden_dir):
    """Fixtures made by the real libraries in the reference's own call chain."""
    from b200clip import capi

    g = np.load(os.path.join(golden_dir, "vitb32_1080p.npz"))
    hd = np.concatenate([structured_frames(4, 1080, 1920, seed=7), noise_frames(2, 1080, 1920, seed=8)])
    chw = model_b32.preprocess_u8(torch.from_numpy(hd).cuda(), capi.RESIZE_REFERENCE, chw=True).cpu().numpy()
    for i in range(len(hd)):
        assert np.array_equal(_u8_from_chw(chw[i]), g["pre_u8"][i]), f"1080p frame {i}"
    geo = np.load(os.path.join(golden_dir, "preprocess_geometries.npz"))
    for key in geo.files:
        dims, seed = key.split("_s")
        w, h = (int(v) for v in dims.split("x"))
        f = noise_frames(1, h, w, seed=int(seed))
        got = model_b32.preprocess_u8(torch.from_numpy(f).cuda(), capi.RESIZE_REFERENCE, chw=True)[0].cpu().numpy()
        assert np.array_equal(_u8_from_chw(got), geo[key]), key


def test_strided_frames_and_empty(model_b32):
    """row/frame strides larger than the packed size (a crop of a bigger buffer), and n = 0."""
    from b200clip import capi
    from oracle import preprocess_ref as P

    big = noise_frames(2, 300, 400, seed=21)
    view = big[:, 10:250, 20:340]            # 240 x 320 window
    dev = torch.from_numpy(big).cuda()
    out = torch.empty(2, 3, 224, 224, device="cuda")
    base = dev.data_ptr() + (10 * 400 + 20) * 3
    model_b32.handle.call("b200clip_preprocess_u8_chw", capi._p(base), 2, 240, 320, 300 * 400 * 3, 400 * 3,
                          capi.RESIZE_REFERENCE, capi._p(out), model_b32._stream())
    for i in range(2):
        want = P.to_chw_normalized(P.reference_preprocess_u8(np.ascontiguousarray(view[i])))
        assert np.array_equal(out[i].cpu().numpy(), want)
    model_b32.handle.call("b200clip_preprocess_u8_chw", capi._p(None), 0, 240, 320, 0, 0, capi.RESIZE_REFERENCE,
                          capi._p(None), model_b32._stream())


@pytest.mark.parametrize("pad_w,x_off", [(4, 1), (4, 2), (3, 1), (1, 0)])
def test_1080p_window_alignments(model_b32, pad_w, x_off):
    """The 1080p area stage has two code paths: the strip walker (row stride a multiple of 4 bytes, any start
    alignment) and the per-pixel kernel (everything else).  Both must reproduce cv2 bit for bit."""
    from b200clip import capi
    from oracle import preprocess_ref as P

    Wb = 1920 + pad_w
    big = noise_frames(2, 1083, Wb, seed=31 + pad_w)
    view = big[:, 2:1082, x_off:x_off + 1920]
    dev = torch.from_numpy(big).cuda()
    out = torch.empty(2, 3, 224, 224, device="cuda")
    base = dev.data_ptr() + (2 * Wb + x_off) * 3
    model_b32.handle.call("b200clip_preprocess_u8_chw", capi._p(base), 2, 1080, 1920, 1083 * Wb * 3, Wb * 3,
                          capi.RESIZE_REFERENCE, capi._p(out), model_b32._stream())
    for i in range(2):
        want = P.to_chw_normalized(P.reference_preprocess_u8(np.ascontiguousarray(view[i])))
        assert np.array_equal(out[i].cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("env", [{}, {"B200CLIP_AREA_FP32": "1"}, {"B200CLIP_K1_UNFUSED": "1"},
                                 {"B200CLIP_K1_UNFUSED": "1", "B200CLIP_AREA_NOSTRIP": "1"}, {"B200CLIP_VPASS_GENERIC": "1"}, {"B200CLIP_AREA_HFIRST": "1"},
                                 {"B200CLIP_AREA_HFIRST": "1", "B200CLIP_AREA_PX1": "1"}, {"B200CLIP_K1_PERSISTENT": "1"},
                                 {"B200CLIP_AREA_NOMMA": "1"}, {"B200CLIP_AREA_NOMMA": "1", "B200CLIP_K1_PERSISTENT": "1"},
                                 {"B200CLIP_HPASS_PX1": "1"}, {"B200CLIP_HPASS_PX1": "1", "B200CLIP_K1_UNFUSED": "1"}],
                         ids=["default-imma", "fused-fp32", "unfused-strip",
                              "unfused-per-pixel", "imma-generic-vpass", "int-exact-horizontal-first-2col",
                              "int-exact-horizontal-first-1col", "vertical-first-persistent-ctas",
                              "int-exact-vertical-first", "int-exact-vertical-first-persistent",
                              "generic-hpass-1px", "unfused-generic-hpass-1px"])
def test_k1_code_path_variants(env):
    """Every K1 area-stage implementation (IMMA, integer-exact fused, fp32 fused, strip walker, per-pixel) must give the
    same bytes as cv2; the path is chosen once per process, so each variant runs tests/k1_variant_check.py in its own
    interpreter."""
    import subprocess
    import sys

    e = dict(os.environ)
    for k in ("B200CLIP_AREA_FP32", "B200CLIP_AREA_PX1", "B200CLIP_K1_UNFUSED", "B200CLIP_AREA_NOSTRIP", "B200CLIP_VPASS_GENERIC",
              "B200CLIP_AREA_HFIRST", "B200CLIP_K1_PERSISTENT", "B200CLIP_AREA_NOMMA", "B200CLIP_HPASS_PX1"):
        e.pop(k, None)
    e.update(env)
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "k1_variant_check.py")
    r = subprocess.run([sys.executable, script], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("w,h,mode", [(1920, 1080, "REFERENCE"), (1280, 720, "REFERENCE"), (640, 360, "REFERENCE"),
                                      (224, 224, "REFERENCE"), (300, 300, "BICUBIC"), (1000, 700, "BILINEAR_AA")])
def test_bgr_input_flag_equals_converting_first(model_b32, w, h, mode):
    """capi.INPUT_BGR (B200CLIP_INPUT_BGR): frames in OpenCV's BGR order, as cv2.VideoCapture delivers them -- the
    per-frame cv2.cvtColor(BGR2RGB) of frame_extractor.py:191 folded into K1's final store.  Must be the same BITS as
    converting first, for the patch rows, the fp32 CHW planes (both store kernels) and through the host-frame path."""
    import cv2

    from b200clip import capi

    m = getattr(capi, "RESIZE_" + mode)
    rgb = np.concatenate([noise_frames(1, h, w, seed=w + 7 * h), structured_frames(2, h, w, seed=w + h)])
    bgr = np.stack([cv2.cvtColor(f, cv2.COLOR_RGB2BGR) for f in rgb])
    d_rgb, d_bgr = torch.from_numpy(rgb).cuda(), torch.from_numpy(bgr).cuda()
    for chw in (False, True):
        want = model_b32.preprocess_u8(d_rgb, m, chw=chw)
        got = model_b32.preprocess_u8(d_bgr, m | capi.INPUT_BGR, chw=chw)
        assert torch.equal(want.view(torch.int16 if not chw else torch.int32), got.view(torch.int16 if not chw else torch.int32))
    e_want = model_b32.encode_frames_u8(d_rgb, m, normalize=True)
    e_got = model_b32.encode_frames_u8(d_bgr, m | capi.INPUT_BGR, normalize=True)
    assert torch.equal(e_want, e_got)
    h_got = model_b32.encode_frames_u8_host(bgr, m | capi.INPUT_BGR, normalize=True)
    assert np.array_equal(np.asarray(h_got), e_want.cpu().numpy())


def test_bgr_flag_is_refused_for_nv12(model_b32):
    from b200clip import capi

    nv12 = torch.zeros(1, 1080 * 3 // 2, 1920, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="resize mode"):
        model_b32.preprocess_nv12(nv12, capi.RESIZE_REFERENCE | capi.INPUT_BGR)
