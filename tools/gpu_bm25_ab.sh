#!/bin/bash
# A/B of liblrag variants (tools/build_variant.sh) on the BM25 bench.  usage: tools/gpu_bm25_ab.sh "<bench args>" variant...
ARGS=$1; shift
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f GB/s %.1f frac %.3f ms/step %.2f" % (r["kernel_ms"],r["achieved"],r["frac"],d["ms_per_step"]))
'
for v in "$@"; do
  echo "== $v"
  LRAG_LIB_PATH=$PWD/legal_rag_b200/variants/$v.so timeout 600 python bench.py --workload bm25 --steps 3 --warmup 3 --no-cpu-baseline $ARGS 2> gpurun_out/ab_err.log | python -c "$fmt"
  tail -2 gpurun_out/ab_err.log
done
