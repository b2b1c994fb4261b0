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
