"""Long loop of the hybrid step on one GPU (rare-failure hunt): builds bench.py's hybrid workload, then runs `steps` steps with a
synchronise + progress line every `every` steps.  usage: tools/gpu_stress_hybrid.py <dense_sms|auto> <steps> [every] [mode]
mode: step (device-resident inputs), e2e (host buffers), scans (the two scans only)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dense_sms, steps = sys.argv[1], int(sys.argv[2])
every = int(sys.argv[3]) if len(sys.argv) > 3 else 50
mode = sys.argv[4] if len(sys.argv) > 4 else "step"
args = argparse.Namespace(n_docs=0, dim=0, vocab=0, nq=0, k=100, kc=0, colbert_mode="rerank", dense_sms=dense_sms)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.HybridWorkload(args, 0, 1, dev)
wl.setup()
print("partition", wl.partition["dense_sms"], wl.partition["tried_ms_per_step"], flush=True)
sh = wl.shard
t0 = time.time()
done = 0
try:
    while done < steps:
        for _ in range(every):
            if mode == "e2e":
                wl.e2e_step()
            elif mode == "scans":
                sh._scans_side_by_side(wl.Qd, wl.q_indptr, wl.q_term, wl.mx, wl.kc) if sh.dense_sms > 0 else None
            else:
                wl.step()
        torch.cuda.synchronize()
        done += every
        print(f"{done} steps ok, {time.time() - t0:.1f} s", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"FAILED between step {done} and {done + every}: {e!r}"[:600], flush=True)
    sys.exit(3)
print("all ok", flush=True)
