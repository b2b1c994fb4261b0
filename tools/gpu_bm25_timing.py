#!/usr/bin/env python
"""Debug: per-phase cycle breakdown of the BM25 consumer warps (library built with -DLRAG_BM25_TIMING:
tools/build_variant.sh timing -DLRAG_BM25_TIMING).
usage: LRAG_LIB_PATH=legal_rag_b200/variants/timing.so python tools/gpu_bm25_timing.py [n_docs nq mean_len]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legal_rag_b200 import engine, synth, _native
n, nq, ml = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (12_500_000, 4096, 24.0)
dev = torch.device("cuda", 0)
index, st = synth.bm25_synthetic_index(n, 500_000, 10, dev, mean_len=ml)
qi, qt, mx = synth.bm25_synthetic_queries(nq, 500_000, 11, dev)
lib = _native.init(0)
out = (C.c_ulonglong * 16)()
print("dense rows:", index.build_dense_rows())
for it in range(3):
    lib.lrag_bm25_debug_timing(out, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.bm25_topk(index, qi, qt, mx, 100); e1.record(); torch.cuda.synchronize()
    lib.lrag_bm25_debug_timing(out, 0)
    v = list(out)
    tot = sum(v[:5])
    names = ["table wait", "slab init", "chunk loop", "slab end", "item begin/end"]
    print(f"step {e0.elapsed_time(e1):.2f} ms; consumer warp-cycle share: " + ", ".join(f"{n_} {100 * v[i] / tot:.1f}%" for i, n_ in enumerate(names))
          + f"; tables/warp-sum {v[5]} chunks {v[6]} slabs {v[7] // 16}; cycles per slab: " + ", ".join(f"{n_} {16 * v[i] / max(1, v[7]):.0f}" for i, n_ in enumerate(names))
          + f"; cycles/chunk {v[2] / max(1, v[6]):.0f}; chunks per slab {16 * v[6] / max(1, v[7]):.1f}")
    print("   slab end by path [cold, hot list, full scan, full scan + select]: share of slabs " + ", ".join(f"{v[12 + i] / max(1, v[7]):.3f}" for i in range(4))
          + "; cycles per such slab per warp " + ", ".join(f"{v[8 + i] / max(1, v[12 + i]):.0f}" for i in range(4))
          + "; share of slab-end cycles " + ", ".join(f"{v[8 + i] / max(1, v[3]):.2f}" for i in range(4)))

import numpy as np
tr = (C.c_longlong * (1024 * 16 * 4))()
lib.lrag_bm25_debug_trace(tr)
T = np.frombuffer(tr, dtype=np.int64).reshape(1024, 16, 4)[100:1000]
t0 = T[:, :, 0].min(axis=1, keepdims=True)
beg, end, nch, done = T[:, :, 0] - t0, T[:, :, 1] - t0, T[:, :, 2], T[:, :, 3] - t0
dur = end - beg
print("slab period (cycles): mean %.0f" % np.diff(T[:, 0, 3]).mean())
print("chunk loop begin skew (max - min over warps): mean %.0f" % (beg.max(axis=1) - beg.min(axis=1)).mean())
print("chunk loop duration: mean %.0f, max over warps mean %.0f, min over warps mean %.0f" % (dur.mean(), dur.max(axis=1).mean(), dur.min(axis=1).mean()))
print("chunk loop end: mean %.0f, max over warps mean %.0f" % (end.mean(), end.max(axis=1).mean()))
print("barrier passed: mean %.0f" % done.mean())
w = nch > 0
print("cycles per chunk: mean %.0f  p10 %.0f p50 %.0f p90 %.0f p99 %.0f" % ((dur[w] / nch[w]).mean(), *np.percentile(dur[w] / nch[w], [10, 50, 90, 99])))
for k in range(0, 6):
    m = nch == k
    if m.any(): print(f"warps with {k} chunks: {m.mean():.2f} of all, duration mean {dur[m].mean():.0f} p90 {np.percentile(dur[m], 90):.0f}")
for s_ in (0, 1, 2, 300, 301):
    print("slab", s_, "chunks", nch[s_].tolist(), "dur", dur[s_].tolist(), "done", int(done[s_].max()))
