#!/usr/bin/env python
"""Debug: per-phase cycle breakdown of the BM25 consumer warps (library built with -DLRAG_BM25_TIMING).
usage: LRAG_LIB_PATH=.../timing.so python tools/gpu_bm25_timing.py [n_docs nq mean_len]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legal_rag_b200 import engine, synth, _native
n, nq, ml = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (12_500_000, 4096, 24.0)
dev = torch.device("cuda", 0)
index, st = synth.bm25_synthetic_index(n, 500_000, 10, dev, mean_len=ml)
qi, qt, mx = synth.bm25_synthetic_queries(nq, 500_000, 11, dev)
lib = _native.init(0)
out = (C.c_ulonglong * 8)()
for it in range(3):
    lib.lrag_bm25_debug_timing(out, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.bm25_topk(index, qi, qt, mx, 100); e1.record(); torch.cuda.synchronize()
    lib.lrag_bm25_debug_timing(out, 0)
    v = list(out)
    tot = sum(v[:2]) + v[3] + v[4]
    names = ["queue wait", "pieces", "(first barrier)", "slab end", "item begin/end"]
    print(f"step {e0.elapsed_time(e1):.2f} ms; per consumer warp-cycle share: " + ", ".join(f"{n_} {100 * v[i] / tot:.1f}%" for i, n_ in enumerate(names))
          + f"; chunks {v[5]} pieces {v[6]} cycles/chunk {v[1] / max(1, v[5]):.0f} total warp-cycles {tot:.3e}")

import numpy as np
tr = (C.c_longlong * (512 * 16 * 5))()
lib.lrag_bm25_debug_trace(tr, 1)
torch.cuda.synchronize(); engine.bm25_topk(index, qi, qt, mx, 100); torch.cuda.synchronize()
lib.lrag_bm25_debug_trace(tr, 0)
T = np.frombuffer(tr, dtype=np.int64).reshape(512, 16, 5)
ok = (T[:, :, 4] > 0).all(axis=1)
T = T[ok][20:400]
end_prev = np.concatenate([[T[0, :, 4].min()], T[:-1, :, 4].max(axis=1)])
def rel(x): return x - end_prev[:, None]
has_piece = T[:, :, 0] > 0
print("slabs traced", len(T), "mean slab period", np.diff(T[:, 0, 4]).mean())
print("first piece read  (after prev slab end): mean %.0f  max-over-warps mean %.0f" % (rel(T[:, :, 0])[has_piece].mean(), np.where(has_piece, rel(T[:, :, 0]), 0).max(axis=1).mean()))
print("last piece done   : mean %.0f  max-over-warps mean %.0f" % (rel(T[:, :, 1])[has_piece].mean(), np.where(has_piece, rel(T[:, :, 1]), 0).max(axis=1).mean()))
print("marker read       : mean %.0f  max-over-warps mean %.0f  min-over-warps mean %.0f" % (rel(T[:, :, 2]).mean(), rel(T[:, :, 2]).max(axis=1).mean(), rel(T[:, :, 2]).min(axis=1).mean()))
print("barrier 1 passed  : mean %.0f" % rel(T[:, :, 3]).mean())
print("slab end done     : mean %.0f" % rel(T[:, :, 4]).mean())
print("warps with pieces per slab: %.1f" % has_piece.sum(axis=1).mean())
for s_ in range(3):
    print("slab", s_, "per warp [first piece, last done, marker, bar1, end]:")
    print(rel(T[s_:s_ + 1])[0].astype(int) if False else (T[s_] - end_prev[s_]).astype(int))
