"""Pins oracle/clip_ref.py: (1) against the golden embeddings the reference's own OpenCLIPModel wrapper produced on
top of it (tests/golden/make_golden.py), (2) against an independent CLIP implementation (transformers.CLIPModel)
with the same weights."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_ref
from oracle import preprocess_ref as P
from synth import QUERIES, structured_frames


def test_golden_vitb32_subset(golden_dir, oracle_sd_b32):
    g = np.load(os.path.join(golden_dir, "vitb32_cfg1.npz"))
    ref = clip_ref.CLIPRef(clip_ref.CONFIGS["ViT-B-32"], oracle_sd_b32)
    frames = structured_frames(128, 224, 224, seed=1234)[:6]
    x = torch.from_numpy(np.stack([P.to_chw_normalized(P.clip_transform_u8(f)) for f in frames]))
    e = ref.encode_image(x)
    e = (e / e.norm(dim=-1, keepdim=True)).numpy()
    assert np.abs(e - g["emb"][:6]).max() < 2e-5
    t = ref.encode_text(clip_ref.synthetic_tokenize(list(QUERIES)))
    t = (t / t.norm(dim=-1, keepdim=True)).numpy()
    assert np.abs(t - g["txt"]).max() < 2e-5
    assert np.abs(e @ t.T - g["scores"][:6]).max() < 2e-5


def _hf_from_oracle(cfg, sd):
    tr = pytest.importorskip("transformers")
    hc = tr.CLIPConfig(
        projection_dim=cfg.embed_dim,
        vision_config=dict(hidden_size=cfg.width, intermediate_size=cfg.mlp_dim, num_hidden_layers=cfg.layers,
                           num_attention_heads=cfg.heads, image_size=cfg.image_size, patch_size=cfg.patch,
                           hidden_act="quick_gelu", layer_norm_eps=cfg.ln_eps, projection_dim=cfg.embed_dim),
        text_config=dict(hidden_size=cfg.text_width, intermediate_size=cfg.text_mlp_dim,
                         num_hidden_layers=cfg.text_layers, num_attention_heads=cfg.text_heads,
                         max_position_embeddings=cfg.text_ctx, vocab_size=cfg.text_vocab, hidden_act="quick_gelu",
                         layer_norm_eps=cfg.ln_eps, projection_dim=cfg.embed_dim, eos_token_id=cfg.text_vocab - 1,
                         bos_token_id=cfg.text_vocab - 2, pad_token_id=0))
    m = tr.CLIPModel(hc).eval()
    hs = m.state_dict()

    def put(k, v):
        assert hs[k].shape == v.shape, (k, hs[k].shape, v.shape)
        hs[k] = v.clone()

    put("vision_model.embeddings.class_embedding", sd["visual.class_embedding"])
    put("vision_model.embeddings.position_embedding.weight", sd["visual.positional_embedding"])
    put("vision_model.embeddings.patch_embedding.weight", sd["visual.conv1.weight"])
    put("vision_model.pre_layrnorm.weight", sd["visual.ln_pre.weight"])
    put("vision_model.pre_layrnorm.bias", sd["visual.ln_pre.bias"])
    put("vision_model.post_layernorm.weight", sd["visual.ln_post.weight"])
    put("vision_model.post_layernorm.bias", sd["visual.ln_post.bias"])
    put("visual_projection.weight", sd["visual.proj"].T)
    put("text_model.embeddings.token_embedding.weight", sd["token_embedding.weight"])
    put("text_model.embeddings.position_embedding.weight", sd["positional_embedding"])
    put("text_model.final_layer_norm.weight", sd["ln_final.weight"])
    put("text_model.final_layer_norm.bias", sd["ln_final.bias"])
    put("text_projection.weight", sd["text_projection"].T)
    for tower, src, width, layers in (("vision_model", "visual.transformer", cfg.width, cfg.layers),
                                      ("text_model", "transformer", cfg.text_width, cfg.text_layers)):
        for i in range(layers):
            a, b = f"{tower}.encoder.layers.{i}", f"{src}.resblocks.{i}"
            qw, kw, vw = sd[f"{b}.attn.in_proj_weight"].split(width)
            qb, kb, vb = sd[f"{b}.attn.in_proj_bias"].split(width)
            for n, w_, b_ in (("q", qw, qb), ("k", kw, kb), ("v", vw, vb)):
                put(f"{a}.self_attn.{n}_proj.weight", w_)
                put(f"{a}.self_attn.{n}_proj.bias", b_)
            put(f"{a}.self_attn.out_proj.weight", sd[f"{b}.attn.out_proj.weight"])
            put(f"{a}.self_attn.out_proj.bias", sd[f"{b}.attn.out_proj.bias"])
            put(f"{a}.layer_norm1.weight", sd[f"{b}.ln_1.weight"])
            put(f"{a}.layer_norm1.bias", sd[f"{b}.ln_1.bias"])
            put(f"{a}.layer_norm2.weight", sd[f"{b}.ln_2.weight"])
            put(f"{a}.layer_norm2.bias", sd[f"{b}.ln_2.bias"])
            put(f"{a}.mlp.fc1.weight", sd[f"{b}.mlp.c_fc.weight"])
            put(f"{a}.mlp.fc1.bias", sd[f"{b}.mlp.c_fc.bias"])
            put(f"{a}.mlp.fc2.weight", sd[f"{b}.mlp.c_proj.weight"])
            put(f"{a}.mlp.fc2.bias", sd[f"{b}.mlp.c_proj.bias"])
    m.load_state_dict(hs)
    return m


def _tensor(out):
    return out if isinstance(out, torch.Tensor) else out.pooler_output


def test_matches_independent_hf_clip():
    """Cross-check on the tiny geometry (same code paths as ViT-B/32 and ViT-L/14; runs in seconds)."""
    cfg = clip_ref.CONFIGS["ViT-tiny-test"]
    sd = clip_ref.init_state_dict(cfg, seed=5)
    hf = _hf_from_oracle(cfg, sd)
    ref = clip_ref.CLIPRef(cfg, sd)
    x = torch.randn(3, 3, cfg.image_size, cfg.image_size)
    tok = torch.randint(1, cfg.text_vocab - 3, (4, cfg.text_ctx))
    tok[:, 0] = cfg.text_vocab - 2
    for i, L in enumerate([3, 7, 12, 15]):
        tok[i, L] = cfg.text_vocab - 1
        tok[i, L + 1:] = 0
    with torch.no_grad():
        hi = _tensor(hf.get_image_features(pixel_values=x))
        ht = _tensor(hf.get_text_features(input_ids=tok, attention_mask=torch.ones_like(tok)))
    assert torch.allclose(ref.encode_image(x), hi, atol=2e-5), float((ref.encode_image(x) - hi).abs().max())
    assert torch.allclose(ref.encode_text(tok), ht, atol=2e-5), float((ref.encode_text(tok) - ht).abs().max())


def test_synthetic_tokenizer_framing():
    t = clip_ref.synthetic_tokenize(["a b c", "x " * 100])
    assert t.shape == (2, 77) and t[0, 0] == 49406 and t[0, 4] == 49407 and t[0, 5:].sum() == 0
    assert t[1, 76] == 49407 and (t.argmax(-1) == torch.tensor([4, 76])).all()
