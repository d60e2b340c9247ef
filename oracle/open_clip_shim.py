"""ORACLE (test infrastructure, not product): an `open_clip`-shaped module object backed by the CPU
restatement, so that the reference's own unmodified wrapper classes can run on top of it.

The reference imports `open_clip` at src/models/openclip_model.py:2 and uses exactly:
  open_clip.create_model_and_transforms(name, pretrained=..., device=...) -> (model, _, preprocess)  (:77-81)
  open_clip.get_tokenizer(name) -> callable(list[str]) -> LongTensor[Q, 77]                          (:82)
  model.eval(), model.encode_image(x), model.encode_text(tokens)                                     (:83,177,205-208)
  preprocess(PIL.Image) -> FloatTensor[3,224,224]                                                    (:171,193)
`preprocess` here is the genuine torchvision/Pillow pipeline open_clip's `image_transform` builds.
"""
from __future__ import annotations

import types

import torch

from . import clip_ref
from .preprocess_ref import OPENAI_MEAN, OPENAI_STD

_STATE = {"seed": 0, "gain": 1.0, "state_dicts": {}}


def configure(seed: int = 0, gain: float = 1.0):
    _STATE["seed"], _STATE["gain"] = seed, gain


def get_state_dict(model_name: str):
    key = (model_name, _STATE["seed"], _STATE["gain"])
    if key not in _STATE["state_dicts"]:
        _STATE["state_dicts"][key] = clip_ref.init_state_dict(clip_ref.CONFIGS[model_name], _STATE["seed"],
                                                              _STATE["gain"])
    return _STATE["state_dicts"][key]


def image_transform(size: int = 224):
    import torchvision.transforms as T
    from torchvision.transforms import InterpolationMode

    def _rgb(img):
        return img.convert("RGB")

    return T.Compose([
        T.Resize(size, interpolation=InterpolationMode.BICUBIC),
        T.CenterCrop(size),
        _rgb,
        T.ToTensor(),
        T.Normalize(mean=OPENAI_MEAN, std=OPENAI_STD),
    ])


def create_model_and_transforms(model_name: str, pretrained=None, device=None, **_kw):
    cfg = clip_ref.CONFIGS[model_name]
    model = clip_ref.CLIPRef(cfg, get_state_dict(model_name))
    pre = image_transform(cfg.image_size)
    return model, pre, pre


def get_tokenizer(model_name: str):
    cfg = clip_ref.CONFIGS[model_name]

    def tok(texts):
        return clip_ref.synthetic_tokenize(texts, cfg.text_ctx, cfg.text_vocab)

    return tok


def as_module() -> types.ModuleType:
    m = types.ModuleType("open_clip")
    m.create_model_and_transforms = create_model_and_transforms
    m.get_tokenizer = get_tokenizer
    m.__b200clip_oracle__ = True
    return m
