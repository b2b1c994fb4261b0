#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/dense_ab.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dense" > gpurun_out/test_dense.log 2>&1; echo "dense tests exit $? : $(tail -1 gpurun_out/test_dense.log)"
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("kernel_ms %.2f TF %.1f frac %.3f ms/step %.2f e2e_ms %.2f launches %d clk %s" % (r["kernel_ms"],r["achieved"],r["frac"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["gpu_launches"],d["clocks"]))
    else: print(l.rstrip())
'
for cfg in "LRAG_DENSE_ONE_EPOCH=1" "LRAG_X=0" "LRAG_DENSE_DEBUG=2"; do
  echo "== $cfg 10M" >> gpurun_out/dense_ab.log
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> gpurun_out/dense_ab.log
done
for nq in 1 64 128 1024; do
  echo "== nq=$nq 10M" >> gpurun_out/dense_ab.log
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --nq $nq 2>&1 | python -c "$fmt" >> gpurun_out/dense_ab.log
done
echo "== 1M nq 4096" >> gpurun_out/dense_ab.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --n-docs 1000000 2>&1 | python -c "$fmt" >> gpurun_out/dense_ab.log
cat gpurun_out/dense_ab.log
