#!/bin/bash
# Small-batch (online) latency of the channels: nq = 1 and 16 at the configs' corpus sizes.
mkdir -p gpurun_out; : > gpurun_out/latency.log
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("%s kernel_ms %.3f ms/step %.3f e2e_ms %.3f achieved %.0f %s" % (d["config"]["workload"][:60], r["kernel_ms"], d["ms_per_step"], d["e2e"]["ms_per_step"], r["achieved"], r["unit"]))
'
for nq in 1 16; do
  for wl in bm25 dense; do
    python bench.py --workload $wl --nq $nq --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$fmt" | sed "s/^/nq=$nq /" >> gpurun_out/latency.log
  done
done
cat gpurun_out/latency.log
