"""Kernel list of ONE eager hybrid query over the UCC corpus (run under ncu --metrics gpu__time_duration.sum)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from legal_rag_b200 import engine
wl = bench.UccWorkload(argparse.Namespace(nq=4, k=100), 0, 1, torch.device("cuda", 0))
wl.setup()
docs, V, X, Q, queries = wl._corpus()
for q in range(4):
    Qd = torch.from_numpy(Q[q:q + 1]).cuda().to(torch.bfloat16)
    qi = torch.tensor([0, len(queries[q])], dtype=torch.int64).cuda()
    qt = torch.tensor(queries[q], dtype=torch.int32).cuda()
    d = engine.dense_topk(wl.X, Qd, 100)
    b = engine.bm25_topk(wl.index, qi, qt, len(queries[q]), 100)
    engine.fuse_topk(d, b, None, k=100, method="weighted_sum", w_dense=0.6, w_bm25=0.4)
    torch.cuda.synchronize()
