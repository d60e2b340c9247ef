"""Second 1080p golden ("the 1080p golden is 6 frames"): 24 more decoded 1080x1920 frames
through the reference's own, unmodified classes -- MemoryManager.resize_frame_for_memory (memory_manager.py:299-322)
and OpenCLIPModel.encode_images (openclip_model.py:152-198), imported from /root/reference with the stub recipe of
make_golden.py on the seeded fp32 CLIP restatement -- plus the text embeddings and scores of the three test queries.
Only embeddings and scores are stored (frames are regenerated from their seeds by the test).

  python tests/golden/make_golden_1080p_more.py      # writes tests/golden/vitb32_1080p_more.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from make_golden import import_reference  # noqa: E402

N_STRUCTURED, N_NOISE, SEED_STRUCTURED, SEED_NOISE = 20, 4, 4242, 4343


def frames_1080p():
    from synth import noise_frames, structured_frames

    return np.concatenate([structured_frames(N_STRUCTURED, 1080, 1920, seed=SEED_STRUCTURED),
                           noise_frames(N_NOISE, 1080, 1920, seed=SEED_NOISE)])


def main():
    import torch

    from synth import QUERIES

    torch.set_num_threads(os.cpu_count() or 8)
    import_reference("ViT-B-32")
    import src.services  # noqa: F401  (first: the reference has a services <-> pipeline import cycle)
    from src.models.openclip_model import OpenCLIPModel
    from src.utils.memory_manager import memory_manager

    model = OpenCLIPModel(force_device="cpu")
    assert model.model_loaded
    hd = frames_1080p()
    shrunk = np.stack([memory_manager.resize_frame_for_memory(f, 512, 512) for f in hd])
    assert shrunk.shape[1:] == (288, 512, 3)
    emb = model.encode_images(shrunk)
    txt = model.encode_text(list(QUERIES))
    scores = model.compute_similarity(emb, txt)
    np.savez_compressed(os.path.join(HERE, "vitb32_1080p_more.npz"), emb=emb.astype(np.float32), txt=txt.astype(np.float32),
                        scores=scores.astype(np.float32),
                        shrunk_crc=np.array([int(x.astype(np.uint64).sum()) for x in shrunk]))
    print("wrote vitb32_1080p_more.npz", emb.shape, scores.shape)


if __name__ == "__main__":
    main()
