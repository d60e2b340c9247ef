"""No-GPU checks: the C-ABI library loads and exports exactly the symbols include/b200clip.h declares, the ctypes
signatures cover them, error paths are loud without a device, and nothing in the product package imports oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "advanced-video-event-detection-extraction_b200")


def _declared():
    src = open(os.path.join(ROOT, "include", "b200clip.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200clip_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from b200clip import capi

    names = _declared()
    assert len(names) >= 24
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200clip.h but not exported"
    assert sorted(capi.SIGNATURES) == names, set(names) ^ set(capi.SIGNATURES)
    assert capi.load_library().b200clip_version().startswith(b"b200clip")


def test_create_fails_loudly_without_a_b200():
    import torch

    from b200clip import capi
    from b200clip.model_configs import MODEL_CONFIGS, to_capi_config

    if torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10:
        pytest.skip("a B200 is present")
    with pytest.raises(capi.B200ClipError):
        capi.Handle(to_capi_config(MODEL_CONFIGS["ViT-B-32"]), 0)
    from b200clip.models.openclip_model import OpenCLIPModel

    with pytest.raises(RuntimeError):
        OpenCLIPModel()
    with pytest.raises(RuntimeError):
        OpenCLIPModel(force_device="cpu")


def test_bad_config_is_rejected_before_touching_the_device():
    from b200clip import capi

    lib = capi.load_library()
    h = ctypes.c_void_p()
    bad = capi.Config(224, 32, 700, 12, 12, 3072, 512, 0, 1e-5, 77, 49408, 512, 8, 12, 2048)   # width != heads*64
    assert lib.b200clip_create(ctypes.byref(bad), 0, ctypes.byref(h)) == -2
    assert b"width" in lib.b200clip_last_error(None)
    assert lib.b200clip_create(None, 0, ctypes.byref(h)) == -1


def test_missing_library_raises(tmp_path):
    from b200clip import capi

    with pytest.raises(FileNotFoundError):
        capi.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    for dirpath, _dirs, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"
                assert "from oracle" not in txt.replace("oracle/", ""), f


def test_weights_layout_and_flops():
    from b200clip.model_configs import MODEL_CONFIGS
    from b200clip.weights import expected_shapes, random_state_dict
    from oracle import clip_ref

    for name in ("ViT-B-32", "ViT-L-14", "ViT-tiny-test"):
        cfg = MODEL_CONFIGS[name]
        ours = expected_shapes(cfg)
        theirs = clip_ref.init_state_dict(clip_ref.CONFIGS[name], 0) if name == "ViT-tiny-test" else None
        if theirs is not None:
            assert {k: tuple(v.shape) for k, v in theirs.items() if k != "logit_scale"} == dict(ours)
    sd = random_state_dict(MODEL_CONFIGS["ViT-tiny-test"], 1)
    assert all(tuple(sd[k].shape) == s for k, s in expected_shapes(MODEL_CONFIGS["ViT-tiny-test"]).items())
    assert abs(MODEL_CONFIGS["ViT-B-32"].flops_per_image() / 8.8176e9 - 1) < 1e-3
    assert abs(MODEL_CONFIGS["ViT-L-14"].flops_per_image() / 162.03e9 - 1) < 1e-3
    assert MODEL_CONFIGS["ViT-L-14"].patch_k == 640 and MODEL_CONFIGS["ViT-B-32"].patch_k == 3072


def test_tokenizer_surface():
    import torch

    from b200clip.tokenizer import get_tokenizer
    from oracle.clip_ref import synthetic_tokenize

    tok = get_tokenizer("ViT-B-32")
    t = tok(["a person walking", "red car"])
    assert t.shape == (2, 77) and t.dtype == torch.long
    assert torch.equal(t, synthetic_tokenize(["a person walking", "red car"]))
    assert torch.equal(tok("one two"), tok(["one two"]))


def test_python_constants_equal_the_header():
    """Every `#define B200CLIP_<NAME> <int>` of include/b200clip.h that the ctypes binding mirrors (resize modes, the
    BGR input flag, element types, error codes) must carry the same value in capi -- the header is the contract, the
    binding is a copy."""
    import re

    from b200clip import capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "b200clip.h")).read()
    defs = {m.group(1): int(m.group(2).strip("()"), 0)
            for m in re.finditer(r"^#define B200CLIP_(\w+)\s+(\(?-?(?:0x[0-9a-fA-F]+|\d+)\)?)\s*(?:/\*|$)", text, re.M)}
    for name in ("RESIZE_REFERENCE", "RESIZE_BILINEAR_AA", "RESIZE_BICUBIC", "INPUT_BGR", "F32", "BF16"):
        assert defs[name] == getattr(capi, name), name
    for code, label in capi.ERRORS.items():
        assert defs[label] == code, label
    assert defs["OK"] == 0 and defs["INPUT_BGR"] & 0xff == 0      # the flag lives above the resize-mode byte
