"""BASELINE config 3 at full size: ViT-L/14 over a 10-minute 30 fps video (18 000 frames of 224x224) sharded over the
ranks of one box, per-rank top-k merged through ONE NCCL all-gather (b200clip.distributed) -- timed on the device (max
over ranks) and checked for the size-independent property the sharding must keep: the merged top-k (scores, global
indices, intervals) is bit-identical to the top-k rank 0 computes alone over all 18 000 frames.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
           tools/run_config3.py [--frames 18000] [--steps 2]
(N = 1 works too and skips the collective.)"""
import argparse
import json
import os
os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from b200clip import capi
from b200clip import open_clip as oc
from b200clip.distributed import shard_range
from b200clip.model_configs import MODEL_CONFIGS
from b200clip.tokenizer import get_tokenizer
from b200clip.weights import random_state_dict

BLOCK = 250   # frames are generated in global blocks of 250 so that every rank can rebuild any part of the video


def make_frames(lo: int, hi: int, dev) -> torch.Tensor:
    """Frames [lo, hi) of the synthetic video: solid colour + coarse blocks + noise, a function of the GLOBAL index."""
    out = torch.empty(hi - lo, 224, 224, 3, dtype=torch.uint8, device=dev)
    b0 = lo // BLOCK
    while b0 * BLOCK < hi:
        g = torch.Generator(device=dev).manual_seed(7000 + b0)
        f = torch.randint(0, 64, (BLOCK, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
        f += torch.randint(0, 192, (BLOCK, 1, 1, 3), dtype=torch.uint8, device=dev, generator=g)
        blk = torch.randint(0, 192, (BLOCK, 7, 7, 3), dtype=torch.uint8, device=dev, generator=g)
        f //= 2
        f += blk.repeat_interleave(32, 1).repeat_interleave(32, 2) // 2
        s0, s1 = max(lo, b0 * BLOCK), min(hi, (b0 + 1) * BLOCK)
        out[s0 - lo:s1 - lo] = f[s0 - b0 * BLOCK:s1 - b0 * BLOCK]
        b0 += 1
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=18000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--top-k", type=int, default=5)
    ap.add_argument("--model", default="ViT-L-14")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = MODEL_CONFIGS[a.model]
    sd = random_state_dict(cfg, 0)
    model, _, _ = oc.create_model_and_transforms(a.model, state_dict=sd, device=dev, max_images=1024, max_texts=1)
    n_total, k, thr = a.frames, a.top_k, -1.0
    lo, hi = shard_range(n_total, rank, world)
    frames = make_frames(lo, hi, dev)
    tok = get_tokenizer(a.model)(["a person walking across the street"]).to(dev)
    ts = torch.arange(n_total, dtype=torch.float64, device=dev) / 30.0
    dur = n_total / 30.0

    def step():
        txt = model.encode_text(tok, normalize=True)
        emb = model.encode_frames_u8(frames, capi.RESIZE_REFERENCE, normalize=True)
        s, i, iv, c = model.sim_topk_sharded(emb, txt, k, thr, ts, index_base=lo, clip_duration=30.0, video_duration=dur)
        return txt, s, i, iv, c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    step()
    barrier()
    model.handle.reset_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        txt, s, i, iv, c = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = model.handle.launches

    ok = None
    if rank == 0:
        # the whole video on one GPU (same weights, same frames): the sharded result must be the same bits
        del frames
        emb_all = torch.empty(n_total, cfg.embed_dim, device=dev)
        for b0 in range(0, n_total, 3000):
            b1 = min(n_total, b0 + 3000)
            emb_all[b0:b1] = model.encode_frames_u8(make_frames(b0, b1, dev), capi.RESIZE_REFERENCE, normalize=True)
        s1, i1, iv1, c1 = model.sim_topk(emb_all, txt, k, thr, ts, index_base=0, clip_duration=30.0, video_duration=dur)
        torch.cuda.synchronize(dev)
        ok = bool(torch.equal(i, i1) and torch.equal(s.view(torch.int32), s1.view(torch.int32)) and
                  torch.equal(iv, iv1) and torch.equal(c, c1))
        print(json.dumps({
            "config": f"BASELINE configs[2]: {a.model} over a 10-min 30 fps video ({n_total} frames 224x224 uint8), "
                      f"contiguous shards over {world} B200, one NCCL all-gather of top-{k} candidates + merge",
            "n_gpus": world, "frames_total": n_total, "frames_per_gpu": hi - lo, "steps": a.steps,
            "ms_per_step": round(float(ms.item()), 3), "frames_per_s": round(n_total / float(ms.item()) * 1e3, 1),
            "gpu_launches_rank0": int(launches), "top_k_idx": i[0].tolist(), "top_k_scores": [round(v, 6) for v in s[0].tolist()],
            "sharded_topk_bit_identical_to_single_gpu": ok}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok in (None, True) else 1


if __name__ == "__main__":
    sys.exit(main())
