import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the suite runs seeded random weights + the stand-in tokenizer: an explicit opt-in (the product refuses by default)
os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no sm_100 GPU in this process")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle_sd_b32():
    from oracle import clip_ref

    return clip_ref.init_state_dict(clip_ref.CONFIGS["ViT-B-32"], seed=0, gain=1.0)


@pytest.fixture(scope="session")
def model_b32(oracle_sd_b32):
    """ViT-B/32 on the GPU with the oracle's seeded weights (the fixtures in tests/golden were made with them)."""
    from b200clip import open_clip as oc

    model, _, pre = oc.create_model_and_transforms("ViT-B-32", state_dict=oracle_sd_b32, device="cuda:0",
                                                   max_images=256, max_texts=8)
    return model


@pytest.fixture(scope="session")
def tiny_pair():
    """(oracle CLIPRef, B200CLIP) for the tiny test geometry with shared weights."""
    from b200clip import open_clip as oc
    from oracle import clip_ref

    cfg = clip_ref.CONFIGS["ViT-tiny-test"]
    sd = clip_ref.init_state_dict(cfg, seed=3, gain=1.0)
    ref = clip_ref.CLIPRef(cfg, sd)
    model, _, _ = oc.create_model_and_transforms("ViT-tiny-test", state_dict=sd, device="cuda:0")
    return ref, model
