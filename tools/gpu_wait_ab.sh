#!/bin/bash
# A/B of the pipeline waits (suspend-hint vs plain polling in the TMA / MMA warps) for dense and MaxSim.
mkdir -p gpurun_out; : > gpurun_out/wait_ab.log
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("%s kernel_ms %.3f ms/step %.3f achieved %.0f %s" % (d["config"]["workload"][:48], r["kernel_ms"], d["ms_per_step"], r["achieved"], r["unit"]))
'
for variant in hint spin; do
  if [ $variant = spin ]; then LRAG_NVCC_EXTRA="-DLRAG_PIPE_SPIN=1" python -m legal_rag_b200.build --force > /dev/null; fi
  for args in "--workload dense --nq 16 --steps 20" "--workload dense --nq 64 --steps 20" "--workload dense --steps 5" "--workload maxsim --steps 10"; do
    python bench.py $args --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$fmt" | sed "s/^/$variant [$args] /" >> gpurun_out/wait_ab.log
  done
done
cat gpurun_out/wait_ab.log
