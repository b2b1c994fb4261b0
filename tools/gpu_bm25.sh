#!/bin/bash
# BM25 bench sweep.  usage: tools/gpu_bm25.sh "<item slabs for 5M>" "<item slabs for 50M>"
mkdir -p gpurun_out; rm -f gpurun_out/bm25.log
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f GB/s %.1f frac %.3f ms/step %.2f q/s %.0f e2e_ms %.2f postings %s clk %s" % (r["kernel_ms"],r["achieved"],r["frac"],d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["config"]["postings"],d["clocks"]))
    else: print(l.rstrip())
'
run() {
  echo "== slabs=$1 $2" >> gpurun_out/bm25.log
  SECONDS=0
  LRAG_BM25_ITEM_SLABS=$1 timeout 300 python bench.py --workload bm25 --steps 3 --warmup 3 --no-cpu-baseline $2 2> gpurun_out/bm25_err.log | python -c "$fmt" >> gpurun_out/bm25.log
  echo "wall ${SECONDS}s" >> gpurun_out/bm25.log; tail -5 gpurun_out/bm25_err.log >> gpurun_out/bm25.log
}
for s in ${1:-8}; do run $s "--n-docs 5000000 --nq 1024"; done
for s in ${2:-8}; do run $s "--n-docs 50000000 --nq 8192"; done
for s in ${3:-}; do run $s "--n-docs 12500000 --nq 4096 --mean-len 24"; done
cat gpurun_out/bm25.log
