"""configs[0] online shape: latency of one hybrid query (dense + BM25 + fusion over the UCC corpus), eager calls vs the
captured CUDA graph, host vectors in, host hits out."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from legal_rag_b200 import engine
wl = bench.UccWorkload(argparse.Namespace(nq=256, k=100), 0, 1, torch.device("cuda", 0))
wl.setup()
docs, V, X, Q, queries = wl._corpus()
Qh = torch.from_numpy(Q).to(torch.bfloat16).pin_memory()
g = engine.GraphedHybridQuery(wl.X, wl.index, k=100, nq=1, max_terms=32)

def eager(q):
    Qd = Qh[q:q + 1].to("cuda", non_blocking=True)
    qi = torch.tensor([0, len(queries[q])], dtype=torch.int64).pin_memory().to("cuda", non_blocking=True)
    qt = torch.tensor(queries[q], dtype=torch.int32).pin_memory().to("cuda", non_blocking=True)
    d = engine.dense_topk(wl.X, Qd, 100)
    b = engine.bm25_topk(wl.index, qi, qt, len(queries[q]), 100)
    s, i = engine.fuse_topk(d, b, None, k=100, method="weighted_sum", w_dense=0.6, w_bm25=0.4)
    return s.cpu(), i.cpu()

for name, fn in (("eager", eager), ("graph", lambda q: g.search(Qh[q:q + 1], [queries[q]]))):
    for q in range(16):
        fn(q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for q in range(256):
        fn(q)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 256 * 1e6:.1f} us per query (wall clock, host in -> host out, 591 docs, k=100)", flush=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(200):
    g.graph.replay()
ev1.record(); torch.cuda.synchronize()
print(f"graph replay alone: {ev0.elapsed_time(ev1) / 200 * 1e3:.1f} us on the device")
