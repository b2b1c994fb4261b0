#!/usr/bin/env python
"""Hot-spot view of an ncu report: key counters + the SASS lines that hold the most stall samples.
usage: tools/ncu_hot.py <prof.ncu-rep> [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'derived__lts__lts2xbar_bytes.sum.per_second', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'gpc__cycles_elapsed.max',
        'launch__registers_per_thread', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for h, u, v in zip(hdr, units, vals):
    if h in want or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and float(v.replace(',', '') or 0) > 0):
        print(f"{h} [{u}] {v}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()))
h2 = src[1]; ix = {h: i for i, h in enumerate(h2)}; data = src[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data); toti = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
print('total samples', tot, 'total warp-inst', toti)
for n, r in enumerate(data):
    sm = int(r[ix['# Samples']] or 0); ie = int(r[ix['Instructions Executed']] or 0)
    if sm > tot * minpct / 100:
        st = {k: int(r[ix[k]] or 0) for k in h2 if k.startswith('stall_') and 'Not' not in k}
        st = sorted(st.items(), key=lambda x: -x[1])[:2]
        print(n, r[ix['Address']][-5:], f"{100*sm/tot:.1f}%", ie, r[ix['Source']][:80], st)
