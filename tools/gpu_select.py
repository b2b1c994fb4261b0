"""Top-k select over materialised score rows: time and effective bandwidth (4 N bytes per row)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legal_rag_b200 import engine
SHAPES = [(64, 1_000_000, 100), (256, 1_000_000, 100), (8, 50_000_000, 100), (4096, 20_000, 100), (64, 1_000_000, 1000)]
if len(sys.argv) == 4:          # one shape (for an ncu capture): nq N k
    SHAPES = [tuple(int(a) for a in sys.argv[1:4])]
for nq, N, k in SHAPES:
    S = torch.randn((nq, N), device="cuda")
    for _ in range(3):
        engine.topk_select(S, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        engine.topk_select(S, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    engine.prof_enable(8)
    engine.topk_select(S, k)
    torch.cuda.synchronize()
    kms = sum(t for name, t in engine.prof_collect() if name == "select")
    engine.prof_enable(0)
    print(f"topk_select nq={nq} N={N} k={k}: {ms:.3f} ms end to end ({nq * N * 4 / ms / 1e6:.0f} GB/s), "
          f"streaming kernel alone {kms:.3f} ms ({nq * N * 4 / kms / 1e6:.0f} GB/s)", flush=True)
    del S
