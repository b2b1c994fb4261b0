#!/bin/bash
# The hybrid step (power-capped: dense scan + BM25 back to back) for several BM25 dense-row density thresholds.
# usage: tools/gpu_hybrid_dense_sweep.sh [tag] [densities...]
TAG=${1:-hyb}; shift
DENS=${@:-"2 0.5 0.3 0.2 0.1"}
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("ms/step %.2f stages %s kernels %s clk %s" % (d["ms_per_step"], {k: round(v,2) for k,v in d["config"]["stages_ms"].items() if isinstance(v, float)}, {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"]))
'
for dens in $DENS; do
  echo "== hybrid step, BM25 dense rows at min density $dens"
  LRAG_BM25_DENSE_MIN_DENSITY=$dens timeout 600 python bench.py --workload hybrid --steps 6 --warmup 3 --no-cpu-baseline --no-side-blocks 2> gpurun_out/hyb_err_$TAG.log | python -c "$fmt"
  tail -2 gpurun_out/hyb_err_$TAG.log
done
