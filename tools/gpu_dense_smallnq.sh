#!/bin/bash
# dense scan at small query batches (HBM-bound regime), with and without the epoch scheme
mkdir -p gpurun_out; : > gpurun_out/dense_smallnq.log
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("kernel_ms %.3f ms/step %.3f e2e_ms %.3f scan %.0f GB/s launches %d" % (r["kernel_ms"], d["ms_per_step"], d["e2e"]["ms_per_step"], r["scan_gbs"], d["gpu_launches"]))
'
for one in 0 1; do
  for nq in 1 8 16 32 64 128 256; do
    if [ $one = 1 ]; then export LRAG_DENSE_ONE_EPOCH=1; else unset LRAG_DENSE_ONE_EPOCH; fi
    python bench.py --workload dense --nq $nq --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$fmt" | sed "s/^/one_epoch=$one nq=$nq /" >> gpurun_out/dense_smallnq.log
  done
done
cat gpurun_out/dense_smallnq.log
