#!/bin/bash
# BM25 parity tests, then the scan at the hybrid shape for several dense-row density thresholds, then configs[3].
# usage: tools/gpu_bm25_dense.sh [tag] [densities...]
TAG=${1:-dense}; shift
DENS=${@:-"2 0.5 0.3 0.2 0.1"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -m gpu -x -q -k "bm25" > gpurun_out/test_bm25_$TAG.log 2>&1
echo "tests exit $? : $(tail -1 gpurun_out/test_bm25_$TAG.log)"
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f GB/s %.1f frac %.3f ms/step %.2f q/s %.0f e2e_ms %.2f clk %s" % (r["kernel_ms"],r["achieved"],r["frac"],d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["clocks"]["sm_mhz"]))
    else: print(l.rstrip())
'
for dens in $DENS; do
  echo "== hybrid shape, min density $dens"
  LRAG_BM25_DENSE_MIN_DENSITY=$dens timeout 600 python bench.py --workload bm25 --steps 3 --warmup 3 --no-cpu-baseline --n-docs 12500000 --nq 4096 --mean-len 24 2> gpurun_out/bm25_err_$TAG.log | python -c "$fmt"
  tail -2 gpurun_out/bm25_err_$TAG.log
done
for dens in ${CFG3_DENS:-0.2}; do
  echo "== configs[3], min density $dens"
  LRAG_BM25_DENSE_MIN_DENSITY=$dens timeout 600 python bench.py --workload bm25 --steps 3 --warmup 3 --no-cpu-baseline --n-docs 50000000 --nq 8192 2> gpurun_out/bm25_err_$TAG.log | python -c "$fmt"
  tail -2 gpurun_out/bm25_err_$TAG.log
done
