"""world_size-2 gloo test of the sharded path's host logic: contiguous doc shards, per-rank local
top-k with GLOBAL ids, all-gather, k-way merge == single-shard result.  The kernels are replaced by the
oracle here (no GPU); the NCCL + kernel version of the same flow is tests/test_gpu_multirank.py."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense as odense


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, d, nq, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from legal_rag_b200 import engine
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    lo, hi = engine.shard_range(N, world, rank)
    D, I = odense.flat_ip_topk(Q, X[lo:hi], k, id_base=lo)                     # local top-k, global ids
    merge = lambda s, i, kk: tuple(torch.from_numpy(a) for a in odense.merge_topk(s.numpy(), i.numpy(), kk))  # noqa: E731
    s, i = engine.allgather_merge(torch.from_numpy(D), torch.from_numpy(I), k, merge=merge)
    np.save(os.path.join(out_dir, f"s{rank}.npy"), s.numpy())
    np.save(os.path.join(out_dir, f"i{rank}.npy"), i.numpy())
    # several channels in one packed all-gather (HybridShard's exchange): same lists as one all-gather per channel
    Q2 = rng.standard_normal((nq, d)).astype(np.float32)
    D2, I2 = odense.flat_ip_topk(Q2, X[lo:hi], k, id_base=lo)
    many = engine.allgather_merge_many([(torch.from_numpy(D), torch.from_numpy(I)), (torch.from_numpy(D2), torch.from_numpy(I2))],
                                       k, merge=merge)
    one = engine.allgather_merge(torch.from_numpy(D2), torch.from_numpy(I2), k, merge=merge)
    assert torch.equal(many[0][0], s) and torch.equal(many[0][1], i)
    assert torch.equal(many[1][0], one[0]) and torch.equal(many[1][1], one[1])
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shard_gather_merge_equals_single_shard(tmp_path):
    N, d, nq, k = 1001, 32, 9, 20
    mp.spawn(_worker, args=(2, _free_port(), N, d, nq, k, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    D, I = odense.flat_ip_topk(Q, X, k)
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"i{r}.npy"), I)
        np.testing.assert_allclose(np.load(tmp_path / f"s{r}.npy"), D, rtol=1e-5)
