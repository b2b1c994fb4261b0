#!/bin/bash
# Clocks and board power (nvidia-smi, sampled during the timed region) of the hybrid step and of its two big stages alone.
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("ms/step %.2f clocks %s kernels %s" % (d["ms_per_step"], d["clocks"], {k: round(v,2) for k,v in d.get("kernel_ms_per_step", {}).items()}))
'
echo "== hybrid"; timeout 600 python bench.py --workload hybrid --steps 20 --warmup 3 --no-cpu-baseline --no-side-blocks 2>/dev/null | python -c "$fmt"
echo "== bm25 at the hybrid shard shape"; timeout 600 python bench.py --workload bm25 --steps 30 --warmup 3 --no-cpu-baseline --n-docs 12500000 --nq 4096 --mean-len 24 2>/dev/null | python -c "$fmt"
echo "== dense 12.5M x 768"; timeout 600 python bench.py --workload dense --steps 30 --warmup 3 --no-cpu-baseline --n-docs 12500000 --dim 768 2>/dev/null | python -c "$fmt"
