#!/bin/bash
# L2 / DRAM counters of the BM25 scan for several library variants.  usage: tools/gpu_bm25_l2.sh "<bench args>" variant...
ARGS=$1; shift
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
for v in "$@"; do
  export LRAG_LIB_PATH=$PWD/legal_rag_b200/variants/$v.so
  python bench.py --workload bm25 --steps 1 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/l2_plain_$v.log 2>&1 &&
  ncu --metrics $M --clock-control none -k regex:bm25_scan -s 2 -c 1 --csv --log-file gpurun_out/l2_$v.csv python bench.py --workload bm25 --steps 1 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/l2_ncu_$v.log 2>&1
  echo "== $v exit $?"; grep -E "bm25_scan" gpurun_out/l2_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
