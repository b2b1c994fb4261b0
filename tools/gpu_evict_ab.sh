#!/bin/bash
# L2 evict-first hint on the dense scan's corpus tiles: hybrid step (side-by-side scans) and the dense scan alone, off / on.
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("ms/step %.2f kernels %s clk %s W %s" % (d["ms_per_step"], {k: round(v,2) for k,v in d.get("kernel_ms_per_step", {}).items()}, d["clocks"]["sm_mhz"], d["clocks"]["power_w"]))
'
for e in 0 1 0 1; do
  export LRAG_DENSE_X_EVICT_FIRST=$e
  echo "== evict_first=$e: hybrid, dense on ${DENSE_SMS:-70} SMs"; timeout 600 python bench.py --workload hybrid --steps 12 --warmup 3 --no-cpu-baseline --no-side-blocks --dense-sms ${DENSE_SMS:-70} 2>/dev/null | python -c "$fmt"
done
for e in 0 1; do
  export LRAG_DENSE_X_EVICT_FIRST=$e
  echo "== evict_first=$e: dense alone (configs[1])"; timeout 600 python bench.py --workload dense --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$fmt"
done
