#!/bin/bash
# BM25 bench at one shape for several item sizes (LRAG_BM25_ITEM_SLABS).  usage: tools/gpu_bm25_items.sh <variant|-> "<bench args>" slabs...
V=$1; ARGS=$2; shift 2
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f GB/s %.1f frac %.3f ms/step %.2f" % (r["kernel_ms"],r["achieved"],r["frac"],d["ms_per_step"]))
'
[ "$V" != "-" ] && export LRAG_LIB_PATH=$PWD/legal_rag_b200/variants/$V.so
for s in "$@"; do
  echo "== item_slabs $s"
  LRAG_BM25_ITEM_SLABS=$s timeout 600 python bench.py --workload bm25 --steps 3 --warmup 3 --no-cpu-baseline $ARGS 2> gpurun_out/items_err.log | python -c "$fmt"
  tail -2 gpurun_out/items_err.log
done
