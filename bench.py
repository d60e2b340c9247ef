#!/usr/bin/env python
"""bench.py -- frames/sec of the phase1_mvp hot path (K1 preprocess -> ViT-B/32 embed -> text score -> top-k) on B200.

Workload (BASELINE.json configs[1]): ViT-B/32 over a 1-hour video at 1 fps = 3600 decoded 1080p uint8 frames per
GPU, 1 text query, top_k = 5, reference-exact resize chain.  A "step" is one pass over the whole video.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm  (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                         the reference's CPU path on the box's host cores

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with the frames already resident in HBM (device timed,
CUDA events, max over ranks); `e2e` = the same metric through the host-buffer API (pinned host frames -> H2D inside
the timed region -> result rows back on the host); `roofline` = the dominant kernel (tcgen05 GEMM) from CUDA events
recorded around every GEMM launch on the launching stream inside the timed region; `cpu_baseline` = the oracle port
of the reference pipeline timed on a bounded sample of the same workload on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("B200CLIP_ALLOW_SYNTHETIC", "1")   # synthetic weights + stand-in tokenizer: explicit opt-in
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "frames/sec ViT-B/32 embed+score"
UNIT = "frames/s"
QUERY = "a person walking across street"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=3600, help="frames per GPU per step")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--top-k", type=int, default=5)
    ap.add_argument("--cpu-frames", type=int, default=256, help="frames in the cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=64, help="frames per step of the --impl reference arm")
    ap.add_argument("--chunk", type=int, default=0, help="frames per pass of the tower (0 = all frames of the step at once)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md 'clocks' line)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def mark(self):
        """Start of the timed region: samples taken before this (nvidia-smi needs ~1 s to start) are dropped."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        lines = self.lines[self.first:] or self.lines[-3:]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_host_frames(n: int, h: int, w: int, seed: int = 7):
    """Host-side synthetic decoded frames: a few distinct structured + noise frames, cycled."""
    import numpy as np

    from synth import noise_frames, structured_frames

    base = np.concatenate([structured_frames(6, h, w, seed=seed), noise_frames(2, h, w, seed=seed + 1)])
    reps = -(-n // len(base))
    return np.concatenate([base] * reps)[:n]


# =============================================================================================== reference arm
def run_reference(args, rank: int):
    if rank != 0:
        return
    import numpy as np
    import torch

    from b200clip.model_configs import MODEL_CONFIGS
    from b200clip.tokenizer import get_tokenizer
    from b200clip.weights import random_state_dict
    from oracle.reference_pipeline import ReferenceCPU

    cores = os.cpu_count() or 1
    ref = ReferenceCPU("ViT-B-32", state_dict=random_state_dict(MODEL_CONFIGS["ViT-B-32"], 0), threads=cores)
    frames = synth_host_frames(args.ref_frames, args.height, args.width)
    tok = get_tokenizer("ViT-B-32")([QUERY])
    ts = [float(i) for i in range(len(frames))]
    for _ in range(args.warmup):
        ref.query(frames, tok, args.top_k, 0.25, ts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.query(frames, tok, args.top_k, 0.25, ts)
    dt = time.perf_counter() - t0
    value = args.steps * len(frames) / dt
    sample = (f"{len(frames)} of {args.frames} frames ({args.height}x{args.width} uint8) per step: cv2 INTER_AREA shrink + "
              f"PIL/torchvision transform + fp32 PyTorch ViT-B/32 (batch 32) + np.dot/argsort, {cores} threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def workload_config(args, world: int) -> dict:
    return {"workload": "BASELINE configs[1]: ViT-B/32 (QuickGELU) over a 1-hour video at 1 fps -- "
                        f"{args.frames} decoded {args.height}x{args.width} uint8 frames per GPU, 1 text query, "
                        f"top_k={args.top_k}, threshold 0.25, reference resize chain (INTER_AREA<=512 -> PIL bicubic -> crop)",
            "frames_per_gpu": args.frames, "frame_hw": [args.height, args.width], "queries": 1, "top_k": args.top_k,
            "sharding": f"frames sharded over {world} rank(s); one all-gather of top-k candidates" if world > 1
            else "single GPU", "cache": "inputs (22.4 GB of frames per GPU) are far larger than the 126 MB L2",
            "weights": "seeded random init (no checkpoint offline)"}


# =============================================================================================== our arm
def run_b200(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    from b200clip import capi
    from b200clip import open_clip as oc
    from b200clip.distributed import allgather_candidates
    from b200clip.model_configs import MODEL_CONFIGS
    from b200clip.tokenizer import get_tokenizer
    from b200clip.weights import random_state_dict

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = MODEL_CONFIGS["ViT-B-32"]
    sd = random_state_dict(cfg, 0)
    n, H, W = args.frames, args.height, args.width
    model, _, _ = oc.create_model_and_transforms("ViT-B-32", state_dict=sd, device=dev,
                                                 max_images=(args.chunk if 0 < args.chunk < n else n), max_texts=1)
    h = model.handle

    # ---- synthetic decoded frames resident in HBM (stand-in for NVDEC output): per-frame solid colour + blocks + noise
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.empty(n, H, W, 3, dtype=torch.uint8, device=dev)
    for i0 in range(0, n, 200):
        i1 = min(n, i0 + 200)
        f = torch.randint(0, 64, (i1 - i0, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
        bg = torch.randint(0, 192, (i1 - i0, 1, 1, 3), dtype=torch.uint8, device=dev, generator=g)
        f += bg
        blk = torch.randint(0, 192, (i1 - i0, H // 120, W // 120, 3), dtype=torch.uint8, device=dev, generator=g)
        f[:, : (H // 120) * 120, : (W // 120) * 120] //= 2
        f[:, : (H // 120) * 120, : (W // 120) * 120] += (blk.repeat_interleave(120, 1).repeat_interleave(120, 2) // 2)
        frames[i0:i1] = f
        del f, blk
    tok = get_tokenizer("ViT-B-32")([QUERY]).to(dev)
    n_total = n * world
    ts = torch.arange(n_total, dtype=torch.float64, device=dev)  # 1 fps
    thr, k = 0.25, args.top_k

    # (Running the text tower on a side stream under the image tower was measured SLOWER -- 39.5 vs 37.3 ms per step:
    # its small grids take SMs away from the persistent, statically scheduled GEMM CTAs -- so the step is serial.)
    def step():
        txt = model.encode_text(tok, normalize=True)
        emb = model.encode_frames_u8(frames, capi.RESIZE_REFERENCE, normalize=True)
        s, i, iv, c = model.sim_topk(emb, txt, k, thr, ts, index_base=rank * n, clip_duration=30.0,
                                     video_duration=float(n_total))
        if world > 1:
            cs, ci = allgather_candidates(s, i)
            s, i, iv, c = model.topk_merge(cs, ci, thr, ts, 30.0, float(n_total))
        return emb, txt, s, i, iv, c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    h.reset_launches()
    h.profile_read(reset=True)
    h.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        time.sleep(1.2 if not sampler.lines else 0.0)   # let nvidia-smi deliver its first sample before timing
        sampler.mark()
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    h.profile_enable(False)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = h.launches
    prof = h.profile_read(reset=True)
    value = n_total * args.steps / (total_ms / 1e3)

    # ---- end to end through the host-buffer API
    e2e = None
    emb_dev_ref, txt_ref, s_ref, i_ref = out[0], out[1], out[2], out[3]
    host_frames = None
    if not args.no_e2e:
        # pinned host frames per rank: the whole batch on one GPU, smaller slices (re-used `calls` times per step) when
        # several ranks share the host's pinned-memory budget (8 ranks x 450 frames x 6.2 MB = 22 GB)
        n_host = n if world == 1 else min(n, 900 if world <= 2 else 450)
        calls = -(-n // n_host)
        host_frames = torch.empty(n_host, H, W, 3, dtype=torch.uint8, pin_memory=True)
        host_frames.copy_(frames[:n_host])
        tok_host = np.ascontiguousarray(tok.cpu().numpy())
        txt_host = np.empty((1, cfg.embed_dim), np.float32)
        emb_e2e = torch.empty(n_host * calls, cfg.embed_dim, device=dev)
        res_host = [torch.empty(1, k, dtype=torch.float32, pin_memory=True), torch.empty(1, k, dtype=torch.int64, pin_memory=True),
                    torch.empty(1, k, 2, dtype=torch.float64, pin_memory=True), torch.empty(1, dtype=torch.int32, pin_memory=True)]
        ts2 = torch.arange(n_host * calls * world, dtype=torch.float64, device=dev)

        def e2e_step():
            h.call("b200clip_encode_text_host", capi._p(tok_host), 1, capi._p(txt_host), 1, model._stream())
            for c in range(calls):
                model.encode_frames_u8_host(host_frames, capi.RESIZE_REFERENCE, True, out=emb_e2e[c * n_host:(c + 1) * n_host])
            txt = torch.from_numpy(txt_host).to(dev, non_blocking=True)
            s, i, iv, cnt = model.sim_topk(emb_e2e, txt, k, thr, ts2, index_base=rank * n_host * calls,
                                           clip_duration=30.0, video_duration=float(n_host * calls * world))
            if world > 1:
                cs, ci = allgather_candidates(s, i)
                s, i, iv, cnt = model.topk_merge(cs, ci, thr, ts2, 30.0, float(n_host * calls * world))
            for dst, src in zip(res_host, (s, i, iv, cnt)):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_step()
        barrier()
        h.transfer_bytes(reset=True)
        t0 = time.perf_counter()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        h2d_lib, d2h_lib = h.transfer_bytes(reset=True)   # counted inside the library, per cudaMemcpy*Async it issued
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_host * calls * world * e2e_steps / float(dt.item()), "unit": UNIT,
               # library uploads (frame windows + tokens) + the text embedding torch re-uploads for K4
               "h2d_bytes_per_step": int(h2d_lib // e2e_steps + txt_host.nbytes),
               "d2h_bytes_per_step": int(d2h_lib // e2e_steps + sum(t.numel() * t.element_size() for t in res_host)),
               "host_frame_bytes_per_step": int(n_host * calls * H * W * 3),
               "upload": "only the columns/rows of each frame that survive the centre crop are copied (strided cudaMemcpy3DAsync)",
               "steps": e2e_steps, "api": "b200clip_encode_text_host + b200clip_encode_frames_u8_host (pinned host frames, "
               f"{calls} call(s) of {n_host} frames) + b200clip_sim_topk, results copied to host"}

    # ---- CPU baseline (rank 0, N = 1 only): oracle port of the reference pipeline on a bounded sample
    cpu = None
    check = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.reference_pipeline import ReferenceCPU

        cores = os.cpu_count() or 1
        ncpu = min(args.cpu_frames, n)
        sample_np = (host_frames[:ncpu].numpy() if host_frames is not None else frames[:ncpu].cpu().numpy())
        ref = ReferenceCPU("ViT-B-32", state_dict=sd, threads=cores)
        ref.encode_images(sample_np[:32], shrink=True)  # warm-up
        t0 = time.perf_counter()
        emb_cpu = ref.encode_images(sample_np, shrink=True)
        txt_cpu = ref.encode_text_tokens(tok.cpu())
        sims_cpu = (emb_cpu @ txt_cpu.T)[:, 0]
        order_cpu = np.argsort(sims_cpu)[::-1][:k]
        dt = time.perf_counter() - t0
        cpu = {"value": ncpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {ncpu} of the {n} frames: cv2 INTER_AREA + PIL/torchvision transform + fp32 PyTorch "
                         f"ViT-B/32 (batch 32) + np.dot/argsort on {cores} threads, {dt:.1f} s"}
        emb_gpu = emb_dev_ref[:ncpu].cpu().numpy()
        cos = (emb_gpu * emb_cpu).sum(-1)
        sims_gpu = (emb_gpu @ txt_ref.cpu().numpy().T)[:, 0]
        check = {"frames": ncpu, "embedding_cosine_min": float(cos.min()),
                 "max_abs_score_err": float(np.abs(sims_gpu - sims_cpu).max())}

    if rank == 0:
        pk = peaks()
        gemm = prof["gemm"]
        tf = gemm["work"] / (gemm["ms"] / 1e3) / 1e12 if gemm["ms"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            # ncu dram__bytes_read.sum + dram__bytes_write.sum of the GEMM launches (profiles/r01j_gemm_ncu_summary.txt),
            # kept as bytes per FLOP and scaled to this run's average launch
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_flop"] * gemm["work"] / max(gemm["launches"], 1)
        kernels = {}
        for name, r in prof.items():
            if r["launches"]:
                # pre_* are sub-ranges of "preprocess" (not additive with it)
                kernels[name] = {"ms_per_step": r["ms"] / args.steps, "launches_per_step": r["launches"] / args.steps,
                                 "share": r["ms"] / total_ms,
                                 ("tflops" if name == "gemm" else "gbs"): (r["work"] / (r["ms"] / 1e3) / (1e12 if name == "gemm" else 1e9))}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all ViT GEMMs of the step)",
                         "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": tf / pk["bf16_tflops_sustained"], "traffic": traffic,
                         "peak_source": f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step); "
                                        f"burst peak {pk['bf16_tflops']}",
                         "flop_per_launch_avg": gemm["work"] / max(gemm["launches"], 1),
                         "launches": gemm["launches"], "share_of_step": gemm["ms"] / total_ms},
            "kernels": kernels,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "check": check,
            "model_flop_per_frame": cfg.flops_per_image(),
            "model_tflops": value * cfg.flops_per_image() / 1e12 / world,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints to fd 1 (e.g. NCCL's version
    banner) has been redirected to stderr so that it cannot pollute it."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
