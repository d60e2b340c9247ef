"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes shard the frame index space, build local
top-k candidates with GLOBAL indices, exchange them with ONE all-gather and merge.  The merge used here is the
oracle's (the product's merge is the CUDA kernel b200clip_topk_merge, covered by tests/test_gpu_topk.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q, k, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from b200clip.distributed import allgather_candidates, shard_range, sharded_topk
    from oracle import phase1_ref as R

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)                       # same data on every rank
    scores = rng.standard_normal((q, n)).astype(np.float32)
    scores[:, n // 2: n // 2 + 3] = scores[:, 3:4]        # ties across the shard boundary

    calls = []

    def local_topk(lo, hi):
        calls.append((lo, hi))
        s = np.full((q, k), -np.inf, np.float32)
        i = np.full((q, k), -1, np.int64)
        for qq in range(q):
            loc = scores[qq, lo:hi]
            o = np.lexsort((np.arange(lo, hi), loc))[::-1][:k]
            s[qq, :len(o)] = loc[o]
            i[qq, :len(o)] = o + lo
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge(cs, ci):
        ms = np.stack([R.merge_topk_lists(cs[:, qq].numpy(), ci[:, qq].numpy(), k)[0] for qq in range(q)])
        mi = np.stack([R.merge_topk_lists(cs[:, qq].numpy(), ci[:, qq].numpy(), k)[1] for qq in range(q)])
        return ms, mi

    ms, mi = sharded_topk(local_topk, n, merge)
    assert calls == [shard_range(n, rank, world)]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ms=ms, mi=mi, scores=scores)
    # the gather itself: shapes and rank order
    cs, ci = allgather_candidates(*local_topk(*shard_range(n, rank, world)))
    assert cs.shape == (world, q, k) and ci.dtype == torch.int64
    lo1, hi1 = shard_range(n, 1, world)
    valid = ci[1][ci[1] >= 0]
    assert int(valid.min()) >= lo1 and int(valid.max()) < hi1 and valid.numel() == q * min(k, hi1 - lo1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,q,k", [(101, 3, 5), (7, 1, 5)])
def test_two_rank_sharded_topk_equals_global(tmp_path, n, q, k):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, q, k, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["mi"], r1["mi"]) and np.array_equal(r0["ms"], r1["ms"])   # every rank agrees
    for qq in range(q):
        want = np.lexsort((np.arange(n), r0["scores"][qq]))[::-1][:k]
        assert np.array_equal(r0["mi"][qq][:len(want)], want)


def test_shard_range_partitions():
    from b200clip.distributed import shard_range

    for n in (0, 1, 7, 8, 63, 3600, 18000):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) <= -(-n // world) if n else True
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
