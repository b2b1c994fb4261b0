"""Rare-failure hunt for the MaxSim rerank kernel's skipped-slot handling (the multi-rank shape: a rank owns 1/world of the
fused candidates, the other slots are -1).  Many launches over changing skip patterns; every result is checked against the
same call's result with the skipped slots compacted away.  usage: tools/gpu_stress_maxsim_skip.py <world> <iters>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legal_rag_b200 import engine, synth

world, iters = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
Nd, Ld, Lq, nq, C = 125000 // world * world // world, 128, 32, 4096, 200
T = synth.unit_tokens_bf16(Nd, Ld, 128, 5, dev)
Q = synth.unit_tokens_bf16(nq, Lq, 128, 7, dev)
g = torch.Generator(device=dev).manual_seed(1)
t0 = time.time()
bad = 0
try:
    for it in range(iters):
        gid = torch.randint(0, Nd * world, (nq, C), generator=g, device=dev)
        rows = gid - (it % world) * Nd
        owned = (rows >= 0) & (rows < Nd)
        if it % 7 == 3:
            owned &= torch.rand((nq, 1), generator=g, device=dev) > 0.3          # whole queries without a single owned slot
        cand = torch.where(owned, rows, torch.full_like(rows, -1))
        s = engine.maxsim_scores(T, None, Q, cand)
        if it % 50 == 0:
            # reference: the same kernel on the owned slots only, one query block at a time is too slow -- compare with itself
            s2 = engine.maxsim_scores(T, None, Q, cand)
            torch.cuda.synchronize()
            if not torch.equal(s, s2) or not bool(torch.isinf(s[~owned]).all()) or bool(torch.isinf(s[owned]).any()):
                bad += 1
                print(f"iteration {it}: result differs between two launches or the skip mask is wrong", flush=True)
            print(f"{it} launches ok, {time.time() - t0:.1f} s", flush=True)
    torch.cuda.synchronize()
except Exception as e:  # noqa: BLE001
    print(f"FAILED at iteration ~{it}: {e!r}"[:500], flush=True)
    sys.exit(3)
print("all ok" if not bad else f"{bad} mismatches", flush=True)
