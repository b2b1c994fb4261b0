"""Sharded path on real GPUs (needs >= 2): one process per GPU, NCCL all-gather of local top-k lists,
liblrag merge kernel; result must equal the single-GPU search of the whole corpus."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, d, nq, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from legal_rag_b200 import engine, synth
    X = synth.unit_rows_bf16(N, d, 2, "cuda")             # same seed on every rank: the full corpus
    Q = synth.unit_rows_bf16(nq, d, 3, "cuda", chunk=nq)
    lo, hi = engine.shard_range(N, world, rank)
    shard = engine.FlatIPShard(X[lo:hi].contiguous(), id_base=lo)
    s, i = shard.search_device(Q, k)
    if rank == 0:
        fs, fi = engine.dense_topk(X, Q, k)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([bool((fi == i).all()), bool(torch.allclose(fs, s))]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_dense_search_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    mp.spawn(_worker, args=(2, _free_port(), 300_001, 256, 257, 100, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "ok.npy").all()


def _hybrid_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from legal_rag_b200 import engine
    from legal_rag_b200.bm25_index import Bm25HostIndex
    rng = np.random.default_rng(5)                       # same data on every rank: the full corpus
    N, d, V, nq, Ld, Lq, kc, k = 6000, 128, 900, 40, 32, 8, 30, 20
    unit = lambda shape: (lambda x: x / np.linalg.norm(x, axis=-1, keepdims=True))(rng.standard_normal(shape).astype(np.float32))
    X = torch.from_numpy(unit((N, d))).cuda().to(torch.bfloat16)
    Q = torch.from_numpy(unit((nq, d))).cuda().to(torch.bfloat16)
    T = torch.from_numpy(unit((N, Ld, 128))).cuda().to(torch.bfloat16)
    Qt = torch.from_numpy(unit((nq, Lq, 128))).cuda().to(torch.bfloat16)
    docs = [rng.integers(0, V, int(rng.integers(5, 30))) for _ in range(N)]
    host = Bm25HostIndex.from_token_ids(docs, V)
    qi, qt, mx = host.encode_queries([rng.integers(0, V, int(rng.integers(2, 6))).tolist() for _ in range(nq)])
    qi, qt = torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda()
    lo, hi = engine.shard_range(N, world, rank)
    shard = engine.HybridShard(X[lo:hi].contiguous(), host.to_device("cuda", lo, hi), T[lo:hi].contiguous(), None, id_base=lo,
                               tok_row_base=lo, tok_rows_total=N)
    s, i = shard.search_device(Q, qi, qt, mx, Qt, k=k, kc=kc)
    shard.dense_sms = 68             # the two scans side by side on an SM partition: same sharded result
    s2, i2 = shard.search_device(Q, qi, qt, mx, Qt, k=k, kc=kc)
    same_side_by_side = bool(torch.equal(s, s2) and torch.equal(i, i2))
    solo = dist.new_group([0])       # collective call on every rank; the unsharded store below runs in this 1-rank group
    if rank == 0:
        full = engine.HybridShard(X, host.to_device("cuda"), T, None, id_base=0, tok_row_base=0, tok_rows_total=N, group=solo)
        fs, fi = full.search_device(Q, qi, qt, mx, Qt, k=k, kc=kc)
        np.save(os.path.join(out_dir, "hybrid_ok.npy"), np.array([bool((fi == i).all()), bool(torch.allclose(fs, s, rtol=1e-5, atol=1e-6)),
                                                                same_side_by_side]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_hybrid_pipeline_equals_single_gpu(tmp_path):
    """engine.HybridShard over 2 ranks (NCCL all-gather merges + max-reduce of the MaxSim scores) == one unsharded shard."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    mp.spawn(_hybrid_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "hybrid_ok.npy").all()
