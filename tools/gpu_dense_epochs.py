"""Per-launch (epoch) times of the dense scan at small query batches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legal_rag_b200 import engine, synth
N, d = 10_000_000, 1024
X = synth.unit_rows_bf16(N, d, 2, "cuda")
for nq in (1, 16, 32, 33, 48, 64, 96, 128):
    Q = synth.unit_rows_bf16(nq, d, 3, "cuda", chunk=max(nq, 1))
    for _ in range(3):
        engine.dense_topk(X, Q, 100)
    torch.cuda.synchronize()
    engine.prof_enable(64)
    engine.dense_topk(X, Q, 100)
    torch.cuda.synchronize()
    t = [round(ms, 3) for name, ms in engine.prof_collect() if name == "dense_scan"]
    engine.prof_enable(0)
    print(f"nq={nq}: total {sum(t):.3f} ms, epochs {t}", flush=True)
