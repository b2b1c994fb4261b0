#!/bin/bash
# Runs the GPU parity suites one process per kernel family (a faulting kernel poisons its CUDA
# context, so families are isolated), then a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/tests.log
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, -k expression
  echo "=== $1" | tee -a gpurun_out/tests.log
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$2" > gpurun_out/test_$1.log 2>&1
  echo "exit $? : $(tail -1 gpurun_out/test_$1.log)" | tee -a gpurun_out/tests.log
}
run select "topk_select or topk_merge or merge_and_select"
run fuse "fuse"
run bm25 "bm25"
run dense_ref "dense_cuda_core"
run maxsim "maxsim"
run dense "dense_tcgen05 or dense_duplicate or dense_rejects"
echo "=== retrievers" | tee -a gpurun_out/tests.log
timeout 900 python -m pytest tests/test_gpu_retrievers.py -m gpu -x -q > gpurun_out/test_retrievers.log 2>&1
echo "exit $? : $(tail -1 gpurun_out/test_retrievers.log)" | tee -a gpurun_out/tests.log
echo "=== fullsize" | tee -a gpurun_out/tests.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/test_fullsize.log 2>&1
echo "exit $? : $(tail -1 gpurun_out/test_fullsize.log)" | tee -a gpurun_out/tests.log
echo "=== fuzz" | tee -a gpurun_out/tests.log
timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/test_fuzz.log 2>&1
echo "exit $? : $(tail -1 gpurun_out/test_fuzz.log)" | tee -a gpurun_out/tests.log
echo "=== multirank (needs 2 GPUs; skipped on one)" | tee -a gpurun_out/tests.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/test_multirank.log 2>&1
echo "exit $? : $(tail -1 gpurun_out/test_multirank.log)" | tee -a gpurun_out/tests.log
