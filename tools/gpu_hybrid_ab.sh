#!/bin/bash
# Hybrid step (side-by-side scans, dense on $DENSE_SMS SMs) and the BM25 scan alone for library variants ("-" = in-tree build).
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("ms/step %.2f kernels %s clk %s W %s" % (d["ms_per_step"], {k: round(v,2) for k,v in d.get("kernel_ms_per_step", {}).items()}, d["clocks"]["sm_mhz"], d["clocks"]["power_w"]))
'
for v in "$@"; do
  unset LRAG_LIB_PATH
  [ "$v" != "-" ] && export LRAG_LIB_PATH=$PWD/legal_rag_b200/variants/$v.so
  echo "== $v: hybrid"; timeout 600 python bench.py --workload hybrid --steps 12 --warmup 3 --no-cpu-baseline --no-side-blocks --dense-sms ${DENSE_SMS:-68} 2>/dev/null | python -c "$fmt"
  echo "== $v: bm25 alone"; timeout 600 python bench.py --workload bm25 --steps 10 --warmup 3 --no-cpu-baseline --n-docs 12500000 --nq 4096 --mean-len 24 2>/dev/null | python -c "$fmt"
done
