#!/bin/bash
# BM25 tests, then A/B of the bookkeeping warps' poll interval: standalone scan at the hybrid shape and the hybrid step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -m gpu -x -q -k "bm25" 2>&1 | tail -1
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("ms/step %.2f kernels %s clk %s W %s" % (d["ms_per_step"], {k: round(v,2) for k,v in d.get("kernel_ms_per_step", {}).items()}, d["clocks"]["sm_mhz"], d["clocks"]["power_w"]))
'
for v in default poll64 poll1k; do
  if [ $v = default ]; then unset LRAG_LIB_PATH; else export LRAG_LIB_PATH=$PWD/legal_rag_b200/variants/$v.so; fi
  echo "== $v: bm25 alone"; timeout 600 python bench.py --workload bm25 --steps 10 --warmup 3 --no-cpu-baseline --n-docs 12500000 --nq 4096 --mean-len 24 2>/dev/null | python -c "$fmt"
  echo "== $v: hybrid, dense on 68 SMs"; timeout 600 python bench.py --workload hybrid --steps 10 --warmup 3 --no-cpu-baseline --no-side-blocks --dense-sms 68 2>/dev/null | python -c "$fmt"
done
