#!/bin/bash
# BM25 parity tests + bench at the hybrid shape and at configs[3].  usage: tools/gpu_bm25_v2.sh [tag]
TAG=${1:-v2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "bm25" > gpurun_out/test_bm25_$TAG.log 2>&1
echo "tests exit $? : $(tail -1 gpurun_out/test_bm25_$TAG.log)"
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f GB/s %.1f frac %.3f ms/step %.2f q/s %.0f e2e_ms %.2f postings %s clk %s" % (r["kernel_ms"],r["achieved"],r["frac"],d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["config"]["postings"],d["clocks"]))
    else: print(l.rstrip())
'
for cfg in "--n-docs 12500000 --nq 4096 --mean-len 24" "--n-docs 50000000 --nq 8192"; do
  echo "== $cfg"
  timeout 600 python bench.py --workload bm25 --steps 3 --warmup 3 --no-cpu-baseline $cfg 2> gpurun_out/bm25_err_$TAG.log | python -c "$fmt"
  tail -3 gpurun_out/bm25_err_$TAG.log
done
