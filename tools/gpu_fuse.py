"""Fusion kernel: time per launch and effective bandwidth (12 B per candidate in, 12 B per hit out).
usage: tools/gpu_fuse.py [nq kc k]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legal_rag_b200 import engine
SHAPES = [(4096, 100, 100), (4096, 100, 200), (64, 100, 100), (1, 100, 100), (4096, 500, 1000)]
if len(sys.argv) == 4:
    SHAPES = [tuple(int(a) for a in sys.argv[1:4])]
g = torch.Generator(device="cuda"); g.manual_seed(0)
for nq, kc, k in SHAPES:
    chans = []
    for c in range(3):
        s = torch.rand((nq, kc), device="cuda", generator=g).sort(dim=1, descending=True).values * (30 if c == 1 else 1)
        i = torch.stack([torch.randperm(4 * kc, device="cuda", generator=g)[:kc] for _ in range(min(nq, 64))]).repeat((nq + 63) // 64, 1)[:nq]
        chans.append((s.contiguous(), i.contiguous().long()))
    for method in ("weighted_sum", "rrf_norm_blend"):
        for bd in (False, True):
            for _ in range(3):
                engine.fuse_topk(*chans, k=k, method=method, breakdown=bd)
            torch.cuda.synchronize()
            engine.prof_enable(16)
            for _ in range(8):
                engine.fuse_topk(*chans, k=k, method=method, breakdown=bd)
            torch.cuda.synchronize()
            t = [x for name, x in engine.prof_collect() if name == "fuse"]
            engine.prof_enable(0)
            ms = sum(t) / len(t)
            nbytes = nq * (3 * kc * 12 + k * 12 + (k * 32 if bd else 0))
            print(f"fuse nq={nq} kc={kc} k={k} {method} breakdown={bd}: {ms * 1e3:.1f} us ({nbytes / ms / 1e6:.1f} GB/s)", flush=True)
