#!/usr/bin/env python
"""Turns gpurun_out ncu artefacts into the tracked summaries under profiles/.

  tools/ncu_summary.py launches <launches.csv> <out.csv>      # per-launch device times, lrag kernels + share table
  tools/ncu_summary.py raw <prof.ncu-rep> <out.md>            # key counters of every captured launch
"""
import collections
import csv
import subprocess
import sys

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
       "derived__lts__lts2xbar_bytes.sum.per_second", "sm__cycles_active.avg", "gpc__cycles_elapsed.max",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active",
       "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__shared_mem_per_block_dynamic"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    out = [["id", "kernel", "grid", "block", "duration_ms"]]
    for r in rows[1:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ix["Metric Unit"]]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        if "lrag::" in name:
            out.append([r[ix["ID"]], name, r[ix["Grid Size"]], r[ix["Block Size"]], f"{v:.4f}"])
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# share of device time per kernel over the whole command (cold-cache, serialised ncu timings)"])
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow(["#", n[:90], f"launches={a[0]}", f"total_ms={a[1]:.3f}", f"share={a[1] / tot:.4f}"])
        w.writerows(out)


def raw(rep, dst):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full counters, {rep}\n\n")
        names = [r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(f"launch {i} `{n[-28:]}`" for i, n in enumerate(names)) + " |\n")
        f.write("|---|---|" + "---|" * len(names) + "\n")
        for k in KEY:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
