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


def _msg_worker(rank, world, port, q, k, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from b200clip import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7 + rank)
    s = np.sort(rng.standard_normal((q, k)).astype(np.float32), axis=1)[:, ::-1].copy()
    i = (rng.permutation(1000)[: q * k].reshape(q, k) + 1000 * rank).astype(np.int64)
    if rank == 1:
        s[0, k - 1], i[0, k - 1] = -np.inf, -1                      # an empty slot
    msg = D.pack_message(torch.from_numpy(s), torch.from_numpy(i))
    assert msg.numel() == D.msg_bytes(q, k) and msg.numel() % 16 == 0
    buf = torch.empty(world, msg.numel(), dtype=torch.uint8)
    got = D.exchange_messages(msg, out=buf)
    assert got.data_ptr() == buf.data_ptr()                          # the preallocated buffer is used, one collective
    cs, ci = D.unpack_messages(got, q, k)
    assert np.array_equal(cs[rank].numpy(), s) and np.array_equal(ci[rank].numpy(), i)
    ms, mi = D.merge_messages_reference(got.numpy(), q, k)
    np.savez(os.path.join(out_dir, f"msg{rank}.npz"), ms=ms, mi=mi, s=s, i=i)
    assert D.nccl_comm_ptr(torch.device("cpu")) == 0                 # gloo: no NCCL communicator -> generic transport
    dist.barrier()
    dist.destroy_process_group()


def test_packed_message_round_trip_two_ranks(tmp_path):
    """The wire format of the candidate exchange (int64 idx | fp32 score | pad to 16 B) through a real 2-rank
    all_gather_into_tensor, and the merge rule every rank applies to the gathered messages."""
    from b200clip import capi
    from b200clip import distributed as D

    q, k, world = 3, 5, 2
    assert D.msg_bytes(q, k) == capi.load_library().b200clip_topk_msg_bytes(q, k) == 192
    assert D.msg_bytes(1, 1) == 16 and D.msg_bytes(256, 5) == capi.load_library().b200clip_topk_msg_bytes(256, 5)
    mp.spawn(_msg_worker, args=(world, _free_port(), q, k, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "msg0.npz"), np.load(tmp_path / "msg1.npz")
    assert np.array_equal(r0["mi"], r1["mi"]) and np.array_equal(r0["ms"], r1["ms"])
    for qq in range(q):
        cand = sorted([(float(s), int(i)) for r in (r0, r1) for s, i in zip(r["s"][qq], r["i"][qq]) if i >= 0], reverse=True)[:k]
        assert [c[1] for c in cand] == list(r0["mi"][qq]) and np.allclose([c[0] for c in cand], r0["ms"][qq])
