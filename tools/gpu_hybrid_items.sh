#!/bin/bash
# Hybrid step with side-by-side scans for several BM25 item sizes (L2 footprint of the posting ranges all CTAs work on).
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("ms/step %.2f kernels %s clk %s W %s" % (d["ms_per_step"], {k: round(v,2) for k,v in d.get("kernel_ms_per_step", {}).items()}, d["clocks"]["sm_mhz"], d["clocks"]["power_w"]))
'
for s in ${@:-32 16 8}; do
  echo "== item slabs $s"; LRAG_BM25_ITEM_SLABS=$s timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --no-cpu-baseline --no-side-blocks --dense-sms ${DENSE_SMS:-68} 2>/dev/null | python -c "$fmt"
done
