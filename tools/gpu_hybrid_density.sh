#!/bin/bash
# Hybrid step (side-by-side scans, split tuned per run) for several BM25 dense-row density thresholds.
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("ms/step %.2f dense_sms %s kernels %s" % (d["ms_per_step"], d["config"]["sm_partition"]["dense_sms"], {k: round(v,2) for k,v in d.get("kernel_ms_per_step", {}).items()}))
'
for dens in ${@:-0.35 0.5 0.7 2}; do
  echo "== dense rows at min density $dens"; LRAG_BM25_DENSE_MIN_DENSITY=$dens timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --no-cpu-baseline --no-side-blocks 2>/dev/null | python -c "$fmt"
done
