#!/bin/bash
# MaxSim full-scan tests + bench for library variants ("-" = the in-tree build).  usage: tools/gpu_scan_ab.sh variant...
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f %.0f TFLOP/s frac %.3f scan %.0f GB/s" % (r["kernel_ms"], r["achieved"], r["frac"], r.get("scan_gbs", 0)))
'
for v in "$@"; do
  unset LRAG_LIB_PATH
  [ "$v" != "-" ] && export LRAG_LIB_PATH=$PWD/legal_rag_b200/variants/$v.so
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "maxsim" > gpurun_out/test_scan_$v.log 2>&1
  echo "tests: $(tail -1 gpurun_out/test_scan_$v.log)"; grep -E "^E  " gpurun_out/test_scan_$v.log | head -5
  for nq in 64 256; do
    echo -n "nq=$nq "
    timeout 300 python bench.py --workload maxsim_scan --nq $nq --steps 3 --no-cpu-baseline 2>/dev/null | python -c "$fmt"
  done
done
