#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/bm25.log
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f GB/s %.1f frac %.3f ms/step %.2f q/s %.0f e2e_ms %.2f postings %s alg %s clk %s" % (r["kernel_ms"],r["achieved"],r["frac"],d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["config"]["postings"],r["algorithmic"],d["clocks"]))
    else: print(l.rstrip())
'
for args in "--n-docs 5000000 --nq 1024" "--n-docs 50000000 --nq 8192"; do
  echo "== $args" >> gpurun_out/bm25.log
  SECONDS=0
  python bench.py --workload bm25 --steps 3 --warmup 3 --no-cpu-baseline $args 2> gpurun_out/bm25_err.log | python -c "$fmt" >> gpurun_out/bm25.log
  echo "wall ${SECONDS}s" >> gpurun_out/bm25.log; tail -5 gpurun_out/bm25_err.log >> gpurun_out/bm25.log
done
cat gpurun_out/bm25.log
