"""Generates the golden fixtures under tests/golden/ by running the REFERENCE'S OWN, UNMODIFIED classes
(/root/reference/src/models/openclip_model.py::OpenCLIPModel, src/pipeline/phase1_mvp.py::Phase1MVP,
src/services/frame_extractor.py::FrameExtractor.create_sliding_windows, src/utils/memory_manager.py::
resize_frame_for_memory, src/services/video_processor.py::VideoProcessor.preprocess_query) in this container.

Recipe: the third-party modules the reference imports but that are not installed here
(`open_clip`, `ffmpeg`, `mediapipe`, `skimage`) are stubbed in sys.modules; `open_clip` is the oracle shim
(oracle/open_clip_shim.py: CPU fp32 restatement + the genuine torchvision/Pillow transform); DATA_DIR points
at a scratch directory.  Video decode (cv2/decord) is replaced by handing FrameExtractor's output
(frames, timestamps) to Phase1MVP directly -- everything after decode is reference code.

Run:  python tests/golden/make_golden.py        (needs /root/reference; the fixtures are committed)
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REFERENCE = os.environ.get("B200CLIP_REFERENCE", "/root/reference")


def import_reference(model_name: str = "ViT-B-32", seed: int = 0, gain: float = 1.0):
    """Returns the reference's `src` package, imported with the stub recipe."""
    from oracle import open_clip_shim

    open_clip_shim.configure(seed=seed, gain=gain)
    scratch = tempfile.mkdtemp(prefix="b200clip_ref_data_")
    os.environ["DATA_DIR"] = scratch
    os.environ["OPENCLIP_MODEL"] = model_name
    sys.modules["open_clip"] = open_clip_shim.as_module()
    for name in ("ffmpeg", "mediapipe", "skimage", "skimage.metrics"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "skimage.metrics":
                m.structural_similarity = lambda *a, **k: 0.0
            sys.modules[name] = m
    sys.modules["skimage"].metrics = sys.modules["skimage.metrics"]
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    import src.utils.config as config  # noqa: E402

    config.settings.OPENCLIP_MODEL = model_name
    return config


def main():
    import torch

    from synth import QUERIES, noise_frames, structured_frames

    torch.set_num_threads(os.cpu_count() or 8)
    out_dir = HERE

    # ------------------------------------------------------------------ ViT-B/32: config 1
    config = import_reference("ViT-B-32")
    import src.services  # noqa: F401  (first: the reference has a services <-> pipeline import cycle)
    from src.models.openclip_model import OpenCLIPModel
    from src.pipeline.phase1_mvp import Phase1MVP
    from src.services.video_processor import VideoProcessor
    from src.utils.memory_manager import memory_manager

    model = OpenCLIPModel(force_device="cpu")
    assert model.model_loaded and getattr(model.model, "cfg", None) is not None

    frames = structured_frames(128, 224, 224, seed=1234)
    emb = model.encode_images(frames)                      # batch branch (:156-181)
    emb_single = model.encode_images(frames[5])            # single-image branch (:182-198)
    assert np.abs(emb_single[0] - emb[5]).max() < 1e-5
    txt = model.encode_text(list(QUERIES))
    txt_single = model.encode_text(QUERIES[0])
    assert np.abs(txt_single[0] - txt[0]).max() < 1e-6
    scores = model.compute_similarity(emb, txt)
    np.savez_compressed(os.path.join(out_dir, "vitb32_cfg1.npz"), emb=emb.astype(np.float32),
                        txt=txt.astype(np.float32), scores=scores.astype(np.float32))

    # Phase1MVP.process_video on a 512-frame "video" (63 windows), decode replaced by synthetic frames
    config.settings.CONFIDENCE_THRESHOLD = -1.0
    video = structured_frames(512, 224, 224, seed=4321)
    fps = 25.0
    timestamps = [i / fps for i in range(len(video))]
    p1 = Phase1MVP()
    p1.frame_extractor.extract_frames = lambda _path: (video, timestamps)
    vp_query = VideoProcessor.preprocess_query(None, QUERIES[0])
    results = p1.process_video("synthetic.mp4", vp_query, top_k=5)
    results_all, debug = p1.process_video("synthetic.mp4", vp_query, top_k=63, debug_mode=True)
    sims = np.array([d["similarity"] for d in debug], np.float32)
    config.settings.CONFIDENCE_THRESHOLD = float(np.sort(sims)[::-1][2])  # threshold that keeps exactly 3 hits
    results_thr = p1.process_video("synthetic.mp4", vp_query, top_k=5, debug_mode=False)
    thr_used = config.settings.CONFIDENCE_THRESHOLD
    config.settings.CONFIDENCE_THRESHOLD = 0.25
    with open(os.path.join(out_dir, "phase1_cfg1.json"), "w") as f:
        json.dump({
            "query": QUERIES[0], "processed_query": vp_query, "fps": fps, "n_frames": len(video),
            "frames_seed": 4321, "top5": results, "all": results_all, "similarities": sims.tolist(),
            "window_timestamps": [d["timestamp"] for d in debug], "threshold_case": {"threshold": thr_used,
                                                                                      "results": results_thr},
        }, f, indent=1)

    # ------------------------------------------------------------------ ViT-B/32: 1080p chain (config 2)
    hd = np.concatenate([structured_frames(4, 1080, 1920, seed=7), noise_frames(2, 1080, 1920, seed=8)])
    shrunk = np.stack([memory_manager.resize_frame_for_memory(f, 512, 512) for f in hd])
    assert shrunk.shape[1:] == (288, 512, 3)
    emb_hd = model.encode_images(shrunk)
    # the uint8 image the reference's PIL transform produces (before ToTensor), for byte-exact K1 tests
    from PIL import Image

    pre = model.preprocess
    chw = np.stack([pre(Image.fromarray(f)).numpy() for f in shrunk])
    from oracle.preprocess_ref import normalize_table

    tab = normalize_table()
    u8 = np.stack([np.stack([np.abs(tab[c][None, None, :] - chw[i, c][..., None]).argmin(-1) for c in range(3)], -1)
                   for i in range(len(chw))]).astype(np.uint8)
    np.savez_compressed(os.path.join(out_dir, "vitb32_1080p.npz"), emb=emb_hd.astype(np.float32),
                        pre_u8=u8, shrunk_crc=np.array([int(x.astype(np.uint64).sum()) for x in shrunk]))

    # other geometries through the real preprocess only (byte-exact K1 targets)
    geo = {}
    for (w, h, seed) in [(1280, 720, 11), (640, 480, 12), (300, 300, 13), (288, 512, 14), (399, 224, 15),
                         (1024, 1024, 16), (800, 600, 17)]:
        f = noise_frames(1, h, w, seed=seed)[0]
        s = memory_manager.resize_frame_for_memory(f, 512, 512)
        c = pre(Image.fromarray(s)).numpy()
        geo[f"{w}x{h}_s{seed}"] = np.stack(
            [np.abs(tab[ch][None, None, :] - c[ch][..., None]).argmin(-1) for ch in range(3)], -1).astype(np.uint8)
    np.savez_compressed(os.path.join(out_dir, "preprocess_geometries.npz"), **geo)

    # query normalisation (pure string logic)
    qs = ["A person  walks across the street", "The very fast vehicle crashes", "an individual sits",
          "dark blue automobile hits a pedestrian", "really quite pretty canine jumps", "light green  car"]
    with open(os.path.join(out_dir, "queries.json"), "w") as f:
        json.dump({q: VideoProcessor.preprocess_query(None, q) for q in qs}, f, indent=1)

    # ------------------------------------------------------------------ ViT-L/14 (config 3 geometry)
    config = import_reference("ViT-L-14")
    from src.models.openclip_model import OpenCLIPModel as OpenCLIPModelL

    model_l = OpenCLIPModelL(force_device="cpu")
    assert model_l.model.cfg.name == "ViT-L-14"
    frames_l = structured_frames(6, 224, 224, seed=555)
    emb_l = model_l.encode_images(frames_l)
    txt_l = model_l.encode_text(list(QUERIES))
    np.savez_compressed(os.path.join(out_dir, "vitl14.npz"), emb=emb_l.astype(np.float32),
                        txt=txt_l.astype(np.float32))
    print("golden fixtures written to", out_dir)
    for fn in sorted(os.listdir(out_dir)):
        print(f"  {fn}: {os.path.getsize(os.path.join(out_dir, fn))} bytes")


if __name__ == "__main__":
    main()
