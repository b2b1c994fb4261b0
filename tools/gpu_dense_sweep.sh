#!/bin/bash
# timing decomposition of the dense scan: TMA only / TMA+MMA / full, and a batch-size sweep
mkdir -p gpurun_out
for mode in 1 2 0; do
  echo "== LRAG_DENSE_DEBUG=$mode" >> gpurun_out/dense_sweep.log
  LRAG_DENSE_DEBUG=$mode python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('kernel_ms',r['kernel_ms'],'TF',r['achieved'],'scanGB/s',r['scan_gbs'],'ms/step',d['ms_per_step'],'clk',d['clocks'])
    else: print(l.rstrip())
" >> gpurun_out/dense_sweep.log
done
for nq in 64 128 512 1024; do
  echo "== nq=$nq" >> gpurun_out/dense_sweep.log
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --nq $nq "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('kernel_ms',r['kernel_ms'],'TF',r['achieved'],'scanGB/s',r['scan_gbs'],'ms/step',d['ms_per_step'],'clk',d['clocks'])
    else: print(l.rstrip())
" >> gpurun_out/dense_sweep.log
done
cat gpurun_out/dense_sweep.log
