#!/usr/bin/env python
"""Hybrid step with the dense and BM25 scans one after the other vs side by side on an SM partition (HybridShard.dense_sms):
result equality, ms/step, clocks and board power.  usage: python tools/gpu_overlap.py [steps] [dense_sms ...]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
splits = [int(a) for a in sys.argv[2:]] or [60, 68, 76]
args = types.SimpleNamespace(n_docs=0, dim=0, vocab=0, nq=0, k=100, kc=0, colbert_mode="rerank")
dev = torch.device("cuda", 0)
w = bench.HybridWorkload(args, 0, 1, dev)
w.setup()
print("stages one after the other (ms):", {k: round(v, 2) for k, v in w.stages_ms.items() if isinstance(v, float)}, flush=True)
ref = w.step()
torch.cuda.synchronize()

def timed(tag):
    for _ in range(3):
        w.step()
    torch.cuda.synchronize()
    cs = bench.ClockSampler(0); cs.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = w.step()
    e1.record(); torch.cuda.synchronize()
    ck = cs.stop()
    ms = e0.elapsed_time(e1) / steps
    same = torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    print(f"{tag}: {ms:.2f} ms/step ({w.nq / ms:.1f} k queries/s), result identical to the serial step: {same}; sm {ck['sm_mhz']} MHz "
          f"[{ck['sm_mhz_min']}..{ck['sm_mhz_max_seen']}], power {ck['power_w']} W (max {ck['power_w_max']}), {ck['reasons']}", flush=True)

timed("serial       ")
for D in splits:
    w.shard.dense_sms = D
    timed(f"dense on {D:3d} SMs")
w.shard.dense_sms = 0
timed("serial again ")
