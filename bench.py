#!/usr/bin/env python
"""Benchmark of the hybrid-retrieval hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dense|maxsim|maxsim_scan|bm25|hybrid|ucc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic queries.  At N=1 the default
workload is BASELINE.json configs[1] (dense flat-IP top-100, 10M x 1024 bf16, 4096 queries).  With N
ranks every rank owns its own 10M-row shard (the corpus grows to N x 10M: weak scaling), computes its
local top-k and the lists are merged after one NCCL all-gather.  The default run then adds a "hybrid" object to
the same line: every rank's 1/8 shard of configs[4] (12.5M x 768 dense rows + BM25 + ColBERT rerank + fusion)
through the sharded hybrid pipeline, measured the same way (--no-hybrid-block skips it).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "hybrid queries/sec @k=100 (1/2/4/8 B200); dense/MaxSim scan GB/s vs HBM peak"
UNIT = "queries/s"


def measured_traffic(kernel: str, **shape):
    """DRAM bytes per step of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum over the launches of one
    step, from one `ncu --set full` capture kept under profiles/).  Only returned when the capture was taken at exactly
    this workload shape; otherwise null."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            entry = json.load(f).get(kernel)
        if entry and all(entry["shape"].get(k) == v for k, v in shape.items()):
            return entry["dram_bytes_per_step"]
    except (OSError, ValueError, KeyError):
        pass
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "bf16_tflops_burst": p["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =================================================================================================
# workloads
# =================================================================================================
class DenseWorkload:
    """BASELINE.json configs[1]: dense flat inner-product top-100, 10M x 1024 bf16 per GPU, 4096 queries."""
    name = "dense_flat_ip_top100"
    dtype = "bf16"

    def __init__(self, args, rank, world, device):
        self.N, self.d, self.nq, self.k = args.n_docs or 10_000_000, args.dim or 1024, args.nq or 4096, args.k
        self.rank, self.world, self.device = rank, world, device

    def config(self):
        return {"workload": f"configs[1] dense flat-IP top-{self.k}: {self.N} x {self.d} bf16 rows per GPU, batch {self.nq} queries",
                "n_docs_per_gpu": self.N, "dim": self.d, "batch_queries": self.nq, "k": self.k,
                "corpus_rows_total": self.N * self.world,
                "parallelism": f"doc-sharded x{self.world}, local top-k + NCCL all-gather + k-way merge" if self.world > 1 else "single GPU",
                "l2": f"corpus shard {self.N * self.d * 2 / 1e9:.2f} GB >> 126 MB L2 (no flush needed)",
                "value_definition": "rank-level query scans per second: every rank answers the whole batch against its own shard, "
                                    "value = n_gpus * batch_queries * steps / time (queries/s normalised to one shard's corpus)"}

    def setup(self):
        import torch
        from legal_rag_b200 import engine, synth
        self.torch, self.engine = torch, engine
        self.X = synth.unit_rows_bf16(self.N, self.d, 2 + self.rank, self.device)
        self.Q = synth.unit_rows_bf16(self.nq, self.d, 3, self.device, chunk=self.nq)
        self.shard = engine.FlatIPShard(self.X, id_base=self.rank * self.N)
        self.Q_host = self.Q.cpu().pin_memory()
        self.launches_per_step = 3 + (1 if self.world > 1 else 0)

    def step(self):
        return self.shard.search_device(self.Q, self.k)

    def e2e_step(self):
        return self.shard.search(self.Q_host, self.k)

    def e2e_bytes(self):
        return self.nq * self.d * 2, self.nq * self.k * 12

    def units_per_step(self):
        return self.nq * self.world

    dominant = "dense_scan"

    def roofline(self, kernel_ms, peaks):
        flops = 2.0 * self.nq * self.N * self.d
        ach = flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "traffic": measured_traffic("dense_scan_kernel", N=self.N, d=self.d, nq=self.nq, k=self.k),
                "kernel": "dense_scan_kernel", "kernel_ms": kernel_ms,
                "algorithmic": f"2*nq*N*d = {flops:.3e} FLOP per launch", "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                "scan_gbs": self.N * self.d * 2 / (kernel_ms * 1e-3) / 1e9}

    # ---- CPU leg: oracle flat-IP (numpy fp32 BLAS + exact top-k) on a bounded sample ----
    def cpu_sample(self, budget_s=15.0):
        import numpy as np
        from oracle import dense as odense
        n_s, nq_s = 200_000, 256
        rng = np.random.default_rng(2)
        X = rng.standard_normal((n_s, self.d), dtype=np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        Q = rng.standard_normal((nq_s, self.d), dtype=np.float32)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)

        def run():
            t0 = time.perf_counter()
            odense.flat_ip_topk(Q, X, self.k)
            dt = time.perf_counter() - t0
            # linear in corpus rows: time for the full shard = dt * N / n_s
            return nq_s / (dt * self.N / n_s), dt
        sample = (f"oracle numpy-fp32 flat-IP + exact top-{self.k}: {nq_s} queries x {n_s} of the {self.N} rows per step, "
                  f"extrapolated linearly in rows to the full shard")
        return run, sample


class MaxsimWorkload:
    """BASELINE.json configs[2]: ColBERT MaxSim rerank, docs x 128 tokens x 128-d, 32-token queries, 1000 -> 100."""
    name = "maxsim_rerank"
    dtype = "bf16"
    dominant = "maxsim"

    def __init__(self, args, rank, world, device):
        self.Nd, self.Ld, self.Lq = args.n_docs or 1_000_000, 128, 32
        self.nq, self.C, self.k = args.nq or 1024, 1000, args.k
        self.rank, self.world, self.device = rank, world, device

    def config(self):
        return {"workload": f"configs[2] MaxSim rerank: {self.Nd} docs x {self.Ld} x 128 bf16, {self.nq} queries x {self.Lq} tokens, "
                            f"{self.C} candidates -> top-{self.k}", "l2": "gathered bytes per step 33.5 GB >> L2",
                "parallelism": f"x{self.world} replicas of the token store" if self.world > 1 else "single GPU"}

    def setup(self):
        import torch
        from legal_rag_b200 import engine, synth
        self.torch, self.engine = torch, engine
        self.D = synth.unit_tokens_bf16(self.Nd, self.Ld, 128, 5, self.device)
        self.Q = synth.unit_tokens_bf16(self.nq, self.Lq, 128, 7, self.device)
        g = torch.Generator(device=self.device); g.manual_seed(8)
        self.cand = torch.stack([torch.randperm(self.Nd, generator=g, device=self.device)[:self.C] for _ in range(self.nq)]) \
            if self.Nd * self.nq <= 2 ** 26 else torch.randint(0, self.Nd, (self.nq, self.C), generator=g, device=self.device)
        self.Q_host = self.Q.cpu().pin_memory()
        self.cand_host = self.cand.cpu().pin_memory()
        self.launches_per_step = 2

    def step(self):
        return self.engine.maxsim_rerank(self.D, None, self.Q, self.cand, self.k)

    def e2e_step(self):
        q = self.Q_host.to(self.device, non_blocking=True)
        c = self.cand_host.to(self.device, non_blocking=True)
        s, i = self.engine.maxsim_rerank(self.D, None, q, c, self.k)
        return s.cpu(), i.cpu()

    def e2e_bytes(self):
        return self.nq * self.Lq * 128 * 2 + self.nq * self.C * 8, self.nq * self.k * 12

    def units_per_step(self):
        return self.nq * self.world

    def roofline(self, kernel_ms, peaks):
        nbytes = float(self.nq) * self.C * self.Ld * 128 * 2
        ach = nbytes / (kernel_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                "kernel": "maxsim_kernel", "kernel_ms": kernel_ms, "algorithmic": f"nq*C*Ld*dim*2 = {nbytes:.3e} B per launch",
                "peak_source": peaks["source"] + " (copy bandwidth)"}

    def cpu_sample(self, budget_s=15.0):
        import numpy as np
        from oracle import maxsim as omaxsim
        n_s, nq_s = 4000, 8
        rng = np.random.default_rng(5)
        D = rng.standard_normal((n_s, self.Ld, 128), dtype=np.float32)
        Q = rng.standard_normal((nq_s, self.Lq, 128), dtype=np.float32)
        cand = np.stack([rng.permutation(n_s)[:self.C] for _ in range(nq_s)])

        def run():
            t0 = time.perf_counter()
            omaxsim.rerank_topk(Q, D, None, cand, self.k)
            dt = time.perf_counter() - t0
            return nq_s / dt, dt
        return run, f"oracle numpy-fp32 MaxSim: {nq_s} queries x {self.C} candidates x {self.Ld} tokens per step (full per-query work)"


class Bm25Workload:
    """BASELINE.json configs[3]: BM25 over 50M synthetic docs, Zipfian 500k vocab (CSR postings), 8192 multi-term queries."""
    name = "bm25_csr_top100"
    dtype = "f32"
    dominant = "bm25_scan"

    def __init__(self, args, rank, world, device):
        self.N, self.V, self.nq, self.k = args.n_docs or 50_000_000, args.vocab or 500_000, args.nq or 8192, args.k
        self.mean_len = args.mean_len or 40.0
        self.rank, self.world, self.device = rank, world, device

    def config(self):
        return {"workload": f"configs[3] BM25 top-{self.k}: {self.N} docs per GPU (mean length {self.mean_len:g}), Zipf(1) vocab {self.V}, "
                            f"{self.nq} queries of 2-8 terms",
                "postings": getattr(self, "nnz", None), "avgdl": getattr(self, "avgdl", None),
                "l2": "postings streamed per step >> 126 MB L2",
                "parallelism": f"doc-sharded x{self.world}" if self.world > 1 else "single GPU"}

    def setup(self):
        import numpy as np
        import torch
        from legal_rag_b200 import engine, synth
        self.torch, self.engine = torch, engine
        self.index, st = synth.bm25_synthetic_index(self.N, self.V, 10 + self.rank, self.device, id_base=self.rank * self.N,
                                                    mean_len=self.mean_len)
        self.nnz, self.avgdl = st["nnz"], st["avgdl"]
        self.q_indptr, self.q_term, self.mx = synth.bm25_synthetic_queries(self.nq, self.V, 11, self.device)
        df = st["df"].cpu().numpy()
        qi, qt = self.q_indptr.cpu().numpy(), self.q_term.cpu().numpy()
        self.alg_postings = int(sum(int(df[np.unique(qt[qi[j]:qi[j + 1]])].sum()) for j in range(self.nq)))
        self.qi_host, self.qt_host = self.q_indptr.cpu().pin_memory(), self.q_term.cpu().pin_memory()

    def step(self):
        s, i = self.engine.bm25_topk(self.index, self.q_indptr, self.q_term, self.mx, self.k)
        return self.engine.allgather_merge(s, i, self.k)

    def e2e_step(self):
        qi = self.qi_host.to(self.device, non_blocking=True)
        qt = self.qt_host.to(self.device, non_blocking=True)
        s, i = self.engine.bm25_topk(self.index, qi, qt, self.mx, self.k)
        s, i = self.engine.allgather_merge(s, i, self.k)
        return s.cpu(), i.cpu()

    def e2e_bytes(self):
        return self.qi_host.numel() * 8 + self.qt_host.numel() * 4, self.nq * self.k * 12

    def units_per_step(self):
        return self.nq * self.world

    def roofline(self, kernel_ms, peaks):
        nbytes = 8.0 * self.alg_postings
        ach = nbytes / (kernel_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": measured_traffic("bm25_scan_kernel", N=self.N, V=self.V, nq=self.nq, k=self.k, mean_len=self.mean_len),
                "kernel": "bm25_scan_kernel", "kernel_ms": kernel_ms,
                "algorithmic": f"8 B x sum_q sum_(distinct t in q) df(t) = {nbytes:.4e} B per launch", "peak_source": peaks["source"] + " (copy bandwidth)"}

    def cpu_sample(self, budget_s=15.0):
        import numpy as np
        from oracle import bm25 as obm25
        n_s, v_s, nq_s = 20_000, 5_000, 4
        rng = np.random.default_rng(10)
        p = 1.0 / np.arange(1, v_s + 1); p /= p.sum()
        lens = np.clip(np.round(rng.lognormal(np.log(40), 0.6, n_s)), 4, 512).astype(np.int64)
        flat = rng.choice(v_s, size=int(lens.sum()), p=p)
        off = np.concatenate([[0], np.cumsum(lens)])
        lit = obm25.BM25Okapi([[str(t) for t in flat[off[i]:off[i + 1]]] for i in range(n_s)])
        queries = [[str(t) for t in rng.choice(v_s, size=int(rng.integers(2, 9)), p=p)] for _ in range(nq_s)]

        def run():
            t0 = time.perf_counter()
            for q in queries:
                obm25.search(lit, q, self.k)
            dt = time.perf_counter() - t0
            return nq_s / (dt * self.N / n_s), dt
        return run, (f"literal rank_bm25.BM25Okapi.get_scores + Python stable sort (what bm25_retriever.py:74-75 executes, single thread): "
                     f"{nq_s} queries x {n_s} docs per step, extrapolated linearly in docs to {self.N}")

    cpu_cores = 1


class HybridWorkload:
    """BASELINE.json configs[4]: full hybrid (dense 100M x 768 + BM25 + ColBERT rerank + weighted fusion) sharded over the
    GPUs of one box with NCCL top-k merges.  Weak scaling: every rank owns 1/8 of the config's stores (12.5M x 768 dense rows,
    12.5M BM25 docs of mean length 24, 125k x 128 x 128 token rows), so 8 ranks hold exactly configs[4]."""
    name = "hybrid_top100"
    dtype = "bf16"
    dominant = "dense_scan"

    def __init__(self, args, rank, world, device):
        self.N, self.d, self.V = args.n_docs or 12_500_000, args.dim or 768, args.vocab or 500_000
        self.nq, self.k, self.kc = args.nq or 4096, args.k, (getattr(args, "kc", 0) or args.k)
        self.colbert_mode = getattr(args, "colbert_mode", "rerank")
        # rerank: configs[4]'s 1M-doc token store (ids aliased onto it); scan: one token row per document of the shard
        self.Nd_tok, self.Ld, self.Lq = (max(1, self.N // 100), 128, 32) if self.colbert_mode == "rerank" else (self.N, 128, 32)
        self.rank, self.world, self.device = rank, world, device

    def config(self):
        return {"workload": f"configs[4] hybrid top-{self.k}: per GPU {self.N} x {self.d} bf16 dense rows + {self.N} BM25 docs (Zipf vocab {self.V}, "
                            f"mean length 24) + {self.Nd_tok} x {self.Ld} x 128 ColBERT token rows; batch {self.nq} queries; per-channel top-{self.kc}, "
                            + (f"MaxSim over the fused candidate union (<= {2 * self.kc})" if self.colbert_mode == "rerank"
                               else "ColBERT as a first-stage channel: MaxSim of every document (batched full-corpus scan)")
                            + ", weighted_sum 0.6/0.4/0.35",
                "corpus_docs_total": self.N * self.world, "postings": getattr(self, "nnz", None),
                "token_rows": "global id mod total token rows (synthetic aliasing of the id space onto the token store, SURVEY 8d C5)",
                "parallelism": (f"doc-sharded x{self.world}: per channel local top-k + NCCL all-gather + merge, MaxSim by row owner + NCCL max-reduce, "
                                f"replicated fusion") if self.world > 1 else "single GPU",
                "l2": f"dense shard {self.N * self.d * 2 / 1e9:.1f} GB >> 126 MB L2",
                "stages_ms": getattr(self, "stages_ms", None),
                "value_definition": "value = n_gpus * batch_queries * steps / time (every rank answers the whole batch against its shard)"}

    def setup(self):
        import torch
        from legal_rag_b200 import engine, synth
        self.torch, self.engine = torch, engine
        base = self.rank * self.N
        X = synth.unit_rows_bf16(self.N, self.d, 20 + self.rank, self.device)
        index, st = synth.bm25_synthetic_index(self.N, self.V, 10 + self.rank, self.device, mean_len=24.0, id_base=base)
        self.nnz = st["nnz"]
        tokens = synth.unit_tokens_bf16(self.Nd_tok, self.Ld, 128, 5 + self.rank, self.device)
        self.shard = engine.HybridShard(X, index, tokens, None, id_base=base,
                                        tok_row_base=self.rank * self.Nd_tok if self.colbert_mode == "rerank" else base,
                                        tok_rows_total=self.Nd_tok * self.world)
        self.Qd = synth.unit_rows_bf16(self.nq, self.d, 21, self.device, chunk=self.nq)
        self.q_indptr, self.q_term, self.mx = synth.bm25_synthetic_queries(self.nq, self.V, 11, self.device)
        self.Qtok = synth.unit_tokens_bf16(self.nq, self.Lq, 128, 7, self.device)
        self.host = [t.cpu().pin_memory() for t in (self.Qd, self.q_indptr, self.q_term, self.Qtok)]
        self._stage_times()

    def _stage_times(self):
        """Per-stage device times of one step (outside the timed region), reported in config.stages_ms."""
        torch, eng, sh = self.torch, self.engine, self.shard
        ev = lambda: torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            marks = [ev() for _ in range(5)]
            marks[0].record()
            d = eng.allgather_merge(*eng.dense_topk(sh.X, self.Qd, self.kc, sh.id_base), self.kc)
            marks[1].record()
            b = eng.allgather_merge(*eng.bm25_topk(sh.bm25, self.q_indptr, self.q_term, self.mx, self.kc), self.kc)
            marks[2].record()
            if self.colbert_mode == "rerank":
                _, gid = eng.fuse_topk(d, b, None, k=2 * self.kc, method="weighted_sum")
                rows = torch.where(gid >= 0, gid % sh.tok_rows_total, gid) - sh.tok_row_base
                own = (gid >= 0) & (rows >= 0) & (rows < sh.tokens.shape[0])
                eng.maxsim_scores(sh.tokens, None, self.Qtok, torch.where(own, rows, torch.full_like(rows, -1)))
            else:
                eng.allgather_merge(*eng.maxsim_scan_topk(sh.tokens, None, self.Qtok, self.kc, id_base=sh.id_base), self.kc)
            marks[3].record()
            self.step()
            marks[4].record()
            torch.cuda.synchronize()
        self.stages_ms = {"dense+merge": marks[0].elapsed_time(marks[1]), "bm25+merge": marks[1].elapsed_time(marks[2]),
                          "candidates+maxsim": marks[2].elapsed_time(marks[3]), "whole_step": marks[3].elapsed_time(marks[4])}

    def step(self):
        return self.shard.search_device(self.Qd, self.q_indptr, self.q_term, self.mx, self.Qtok, k=self.k, kc=self.kc,
                                        colbert_mode=self.colbert_mode)

    def e2e_step(self):
        return self.shard.search(self.host[0], self.host[1], self.host[2], self.mx, self.host[3], k=self.k, kc=self.kc,
                                 colbert_mode=self.colbert_mode)

    def e2e_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.host), self.nq * self.k * 12

    def units_per_step(self):
        return self.nq * self.world

    def roofline(self, kernel_ms, peaks):
        flops = 2.0 * self.nq * self.N * self.d
        ach = flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "traffic": None, "kernel": "dense_scan_kernel (dominant stage of the hybrid step)", "kernel_ms": kernel_ms,
                "algorithmic": f"2*nq*N*d = {flops:.3e} FLOP per step", "peak_source": peaks["source"] + " (sustained cuBLAS bf16)"}

    def cpu_sample(self, budget_s=15.0):
        """The reference's sequential path (hybrid_retriever.py:282-384) restated: flat-IP + literal BM25 + MaxSim of the
        candidate union + _fuse, single query at a time, on a 20k-doc sample; dense and BM25 extrapolated linearly in docs."""
        import numpy as np
        from oracle import bm25 as obm25, dense as odense, fuse as ofuse, maxsim as omaxsim
        n_s, v_s, nq_s = 20_000, 5_000, 4
        rng = np.random.default_rng(20)
        X = rng.standard_normal((n_s, self.d), dtype=np.float32); X /= np.linalg.norm(X, axis=1, keepdims=True)
        Q = rng.standard_normal((nq_s, self.d), dtype=np.float32); Q /= np.linalg.norm(Q, axis=1, keepdims=True)
        p = 1.0 / np.arange(1, v_s + 1); p /= p.sum()
        lens = np.clip(np.round(rng.lognormal(np.log(24), 0.6, n_s)), 4, 512).astype(np.int64)
        flat = rng.choice(v_s, size=int(lens.sum()), p=p)
        off = np.concatenate([[0], np.cumsum(lens)])
        lit = obm25.BM25Okapi([[str(t) for t in flat[off[i]:off[i + 1]]] for i in range(n_s)])
        queries = [[str(t) for t in rng.choice(v_s, size=int(rng.integers(2, 9)), p=p)] for _ in range(nq_s)]
        D = rng.standard_normal((2000, self.Ld, 128), dtype=np.float32)
        Qt = rng.standard_normal((nq_s, self.Lq, 128), dtype=np.float32)

        def run():
            t_scan = t_rest = 0.0
            for j in range(nq_s):
                t0 = time.perf_counter()
                ds, di = odense.flat_ip_topk(Q[j:j + 1], X, self.kc)
                bs, bi = obm25.search(lit, queries[j], self.kc)
                t1 = time.perf_counter()
                dl = list(zip(di[0].tolist(), ds[0].tolist())); bl = list(zip(bi.tolist(), bs.tolist()))
                cand = np.array([[r["id"] % 2000 for r in ofuse.fuse(dl, bl, [], method="weighted_sum")]])
                cs = omaxsim.maxsim_scores(Qt[j:j + 1], D, None, cand)[0]
                order = np.argsort(-cs, kind="stable")[:self.kc]
                ofuse.fuse(dl, bl, [(int(cand[0][o]), float(cs[o])) for o in order], method="weighted_sum")
                t2 = time.perf_counter()
                t_scan += t1 - t0; t_rest += t2 - t1
            dt = t_scan * self.N / n_s + t_rest
            return nq_s / dt, t_scan + t_rest
        return run, (f"oracle restatement of the reference's sequential hybrid path, one query at a time: numpy flat-IP + literal BM25Okapi over "
                     f"{n_s} docs (extrapolated linearly to {self.N}), MaxSim of the candidate union, _fuse; {nq_s} queries per step")

    cpu_cores = 1


class MaxsimScanWorkload:
    """SURVEY 8d C3, tensor-bound variant: exact full-corpus MaxSim of a query batch against configs[2]'s token store
    (1M docs x 128 tokens x 128-d) -- what the reference's ColBERT channel approximates with PLAID."""
    name = "maxsim_full_scan"
    dtype = "bf16"
    dominant = "maxsim_scan"

    def __init__(self, args, rank, world, device):
        self.Nd, self.Ld, self.Lq = args.n_docs or 1_000_000, 128, 32
        self.nq, self.k = args.nq or 64, args.k
        self.rank, self.world, self.device = rank, world, device

    def config(self):
        return {"workload": f"configs[2] token store, full scan: {self.Nd} docs x {self.Ld} x 128 bf16, batch of {self.nq} queries x {self.Lq} tokens "
                            f"against every document -> top-{self.k}", "l2": f"token store {self.Nd * self.Ld * 256 / 1e9:.1f} GB >> L2",
                "parallelism": f"doc-sharded x{self.world}" if self.world > 1 else "single GPU"}

    def setup(self):
        import torch
        from legal_rag_b200 import engine, synth
        self.torch, self.engine = torch, engine
        self.D = synth.unit_tokens_bf16(self.Nd, self.Ld, 128, 5 + self.rank, self.device)
        self.Q = synth.unit_tokens_bf16(self.nq, self.Lq, 128, 7, self.device)
        self.Q_host = self.Q.cpu().pin_memory()

    def step(self):
        s, i = self.engine.maxsim_scan_topk(self.D, None, self.Q, self.k, id_base=self.rank * self.Nd)
        return self.engine.allgather_merge(s, i, self.k)

    def e2e_step(self):
        q = self.Q_host.to(self.device, non_blocking=True)
        s, i = self.engine.maxsim_scan_topk(self.D, None, q, self.k, id_base=self.rank * self.Nd)
        s, i = self.engine.allgather_merge(s, i, self.k)
        return s.cpu(), i.cpu()

    def e2e_bytes(self):
        return self.nq * self.Lq * 128 * 2, self.nq * self.k * 12

    def units_per_step(self):
        return self.nq * self.world

    def roofline(self, kernel_ms, peaks):
        flops = 2.0 * self.nq * self.Lq * 128 * self.Nd * self.Ld
        ach = flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "traffic": None, "kernel": "maxsim_scan_kernel", "kernel_ms": kernel_ms,
                "algorithmic": f"2*nq*Lq*dim*Nd*Ld = {flops:.3e} FLOP per step", "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                "scan_gbs": self.Nd * self.Ld * 256 / (kernel_ms * 1e-3) / 1e9}

    def cpu_sample(self, budget_s=15.0):
        import numpy as np
        from oracle import maxsim as omaxsim
        n_s, nq_s = 2000, 4
        rng = np.random.default_rng(5)
        D = rng.standard_normal((n_s, self.Ld, 128), dtype=np.float32)
        Q = rng.standard_normal((nq_s, self.Lq, 128), dtype=np.float32)
        cand = np.tile(np.arange(n_s), (nq_s, 1))

        def run():
            t0 = time.perf_counter()
            omaxsim.rerank_topk(Q, D, None, cand, self.k)
            dt = time.perf_counter() - t0
            return nq_s / (dt * self.Nd / n_s), dt
        return run, f"oracle numpy-fp32 MaxSim of {nq_s} queries against {n_s} documents per step, extrapolated linearly in documents to {self.Nd}"


class UccWorkload:
    """BASELINE.json configs[0]: the UCC article corpus (591 docs, vocabulary 3926, 53 991 postings) with random unit
    768-d embeddings, 1024 synthetic queries (random unit vectors + windows of 3-8 consecutive tokens of a random doc), dense + BM25
    top-100 and weighted fusion 0.6/0.4.  Small enough for the literal oracle: this is the CPU-runnable case."""
    name = "ucc_hybrid"
    dtype = "bf16"
    dominant = "dense_scan"

    def __init__(self, args, rank, world, device):
        self.nq, self.k = args.nq or 1024, args.k
        self.rank, self.world, self.device = rank, world, device

    def _corpus(self):
        import numpy as np
        z = np.load(os.path.join(ROOT, "tests", "golden", "ucc_corpus.npz"))
        lens, flat = z["doc_len"].astype(np.int64), z["tokens"].astype(np.int64)
        off = np.concatenate([[0], np.cumsum(lens)])
        docs = [flat[off[i]:off[i + 1]] for i in range(len(lens))]
        X = np.random.default_rng(42).standard_normal((len(docs), 768)).astype(np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        Q = np.random.default_rng(43).standard_normal((self.nq, 768)).astype(np.float32)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)
        rng = np.random.default_rng(44)
        queries = []
        for _ in range(self.nq):
            d = docs[int(rng.integers(0, len(docs)))]
            L = int(rng.integers(3, 9))
            st = int(rng.integers(0, max(1, len(d) - L)))
            queries.append(d[st:st + L].tolist())
        return docs, int(len(z["vocab"])), X, Q, queries

    def config(self):
        return {"workload": f"configs[0] UCC corpus hybrid (dense flat-IP + BM25 + weighted_sum 0.6/0.4), 591 docs, {self.nq} queries, top-{self.k}",
                "l2": "whole corpus fits in L2: this configuration measures launch and host latency, not bandwidth", "parallelism": "single GPU"}

    def setup(self):
        import torch
        from legal_rag_b200 import engine
        from legal_rag_b200.bm25_index import Bm25HostIndex
        self.torch, self.engine = torch, engine
        docs, V, X, Q, queries = self._corpus()
        self.N = len(docs)
        self.kk = min(self.k, self.N)
        self.X = torch.from_numpy(X).to(self.device).to(torch.bfloat16)
        self.Q = torch.from_numpy(Q).to(self.device).to(torch.bfloat16)
        host = Bm25HostIndex.from_token_ids(docs, V)
        self.index = host.to_device(self.device)
        qi, qt, self.mx = host.encode_queries(queries)
        self.qi, self.qt = torch.from_numpy(qi).to(self.device), torch.from_numpy(qt).to(self.device)
        self.host = [self.Q.cpu().pin_memory(), self.qi.cpu().pin_memory(), self.qt.cpu().pin_memory()]

    def _search(self, Q, qi, qt):
        eng = self.engine
        d = eng.dense_topk(self.X, Q, self.kk)
        b = eng.bm25_topk(self.index, qi, qt, self.mx, self.kk)
        return eng.fuse_topk(d, b, None, k=self.kk, method="weighted_sum", w_dense=0.6, w_bm25=0.4)

    def step(self):
        return self._search(self.Q, self.qi, self.qt)

    def e2e_step(self):
        Q, qi, qt = (t.to(self.device, non_blocking=True) for t in self.host)
        s, i = self._search(Q, qi, qt)
        return s.cpu(), i.cpu()

    def e2e_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.host), self.nq * self.kk * 12

    def units_per_step(self):
        return self.nq * self.world

    def roofline(self, kernel_ms, peaks):
        flops = 2.0 * self.nq * self.N * 768
        ach = flops / (max(kernel_ms, 1e-6) * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "traffic": None, "kernel": "dense_scan_kernel (591 docs: three doc tiles; the step is launch-latency bound)", "kernel_ms": kernel_ms,
                "algorithmic": f"2*nq*N*d = {flops:.3e} FLOP per step", "peak_source": peaks["source"] + " (sustained cuBLAS bf16)"}

    def cpu_sample(self, budget_s=15.0):
        """The reference's path in full on this corpus: numpy fp32 flat-IP, literal BM25Okapi.get_scores + stable sort, _fuse."""
        import numpy as np
        from oracle import bm25 as obm25, dense as odense, fuse as ofuse
        docs, V, X, Q, queries = self._corpus()
        lit = obm25.BM25Okapi([[str(t) for t in d] for d in docs])
        nq_s = 64
        kk = min(self.k, len(docs))

        def run():
            t0 = time.perf_counter()
            for j in range(nq_s):
                ds, di = odense.flat_ip_topk(Q[j:j + 1], X, kk)
                bs, bi = obm25.search(lit, [str(t) for t in queries[j]], kk)
                ofuse.fuse(list(zip(di[0].tolist(), ds[0].tolist())), list(zip(bi.tolist(), bs.tolist())), [], method="weighted_sum",
                           w_dense=0.6, w_bm25=0.4)
            dt = time.perf_counter() - t0
            return nq_s / dt, dt
        return run, f"oracle restatement of the reference path on the whole UCC corpus, one query at a time, {nq_s} of the {self.nq} queries per step"

    cpu_cores = 1


WORKLOADS = {"dense": DenseWorkload, "maxsim": MaxsimWorkload, "maxsim_scan": MaxsimScanWorkload, "bm25": Bm25Workload, "hybrid": HybridWorkload,
             "ucc": UccWorkload}


# =================================================================================================
def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port; the third-party wheels it calls are not
    installable here) on the host cores."""
    if rank != 0:
        return
    def timed(name):
        wl = WORKLOADS[name](args, 0, world, None)      # same config object as the native arm at this world size
        run, sample = wl.cpu_sample()
        for _ in range(max(1, min(args.warmup, 2))):
            run()
        vals, t = [], 0.0
        for _ in range(args.steps):
            v, dt = run()
            vals.append(v); t += dt
        value = len(vals) / sum(1.0 / v for v in vals)   # harmonic mean == total queries / total time
        return wl, value, 1e3 * t / args.steps, sample

    # A host answers N shards one after the other: N times the work in N times the time, so the metric's
    # "rank-level query scans per second" does not depend on N for the CPU path.
    wl, value, ms, sample = timed(args.workload)
    cores = getattr(wl, "cpu_cores", os.cpu_count())
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.config(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload == "dense" and not args.no_hybrid_block and not (args.n_docs or args.dim or args.nq):
        wl_h, v_h, ms_h, sample_h = timed("hybrid")      # the counterpart of the native arm's "hybrid" object
        line["hybrid"] = {"value": v_h, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "ms_per_step": ms_h, "config": wl_h.config(),
                          "cpu_baseline": {"value": v_h, "unit": UNIT, "cores": getattr(wl_h, "cpu_cores", os.cpu_count()), "kind": "port",
                                           "sample": sample_h},
                          "e2e": {"value": v_h, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def with_burst(roof, peaks):
    """Tensor-bound rooflines are quoted against the sustained cuBLAS figure (the kernel runs inside a long, power-capped
    step); the same achieved rate against the burst figure is reported next to it -- a stage that follows a cooler stage
    (the dense scan inside the hybrid step) runs at clocks between the two and can exceed the sustained number."""
    if roof.get("bound") == "tensor":
        roof["peak_burst"] = peaks["bf16_tflops_burst"]
        roof["frac_of_burst"] = roof["achieved"] / peaks["bf16_tflops_burst"]
    return roof


def release(wl):
    """Drops a workload's device stores (the CPU leg only needs its shape attributes)."""
    import gc
    import torch
    for name, v in list(vars(wl).items()):
        if not isinstance(v, (int, float, str, bool, dict, type(None), torch.device)) and name not in ("torch", "engine"):
            delattr(wl, name)
    gc.collect()
    torch.cuda.empty_cache()


def measure_native(wl, steps, warmup, rank, world, local_rank, device, peaks):
    """Sets a workload up, times `steps` steps after `warmup` (barrier + synchronize on both sides, CUDA events, max over ranks),
    then the same through the host-buffer API.  Returns the JSON line's fields on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from legal_rag_b200 import engine
    wl.setup()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        wl.step()
    barrier()
    engine.prof_enable(steps * 16 + 8)
    launches0 = engine.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        wl.step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = engine.launch_count() - launches0
    prof = engine.prof_collect()
    engine.prof_enable(0)
    kern = [t for name, t in prof if name == wl.dominant]
    kernel_ms = sum(kern) / steps        # per step (a step may launch the dominant kernel several times)

    # ---- end to end through the host-buffer API: H2D queries -> search -> D2H results, every step ----
    for _ in range(2):
        wl.e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        wl.e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([ms, e2e_ms, kernel_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = t.tolist()
    if rank != 0:
        return None
    units = wl.units_per_step() * steps
    h2d, d2h = wl.e2e_bytes()
    return {"metric": METRIC, "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic (seeded, generated on device; random unit-norm vectors)", "config": wl.config(),
            "clocks": clocks,
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / steps},
            "gpu_launches": launches,
            "roofline": with_burst(wl.roofline(kernel_ms, peaks), peaks),
            "global_queries_per_s": wl.nq * steps / (ms * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="dense", choices=sorted(WORKLOADS))
    ap.add_argument("--n-docs", type=int, default=0)
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--vocab", type=int, default=0)
    ap.add_argument("--nq", type=int, default=0)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--mean-len", type=float, default=0.0, help="bm25: mean document length of the synthetic corpus (default 40)")
    ap.add_argument("--kc", type=int, default=0, help="hybrid: per-channel list length (default k); 500 gives the 1000-candidate rerank of SURVEY 8d C5")
    ap.add_argument("--colbert-mode", default="rerank", choices=["rerank", "scan"],
                    help="hybrid: MaxSim over the fused candidate union, or ColBERT as a first-stage channel over the whole token store")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hybrid-block", action="store_true",
                    help="default workload only: skip the extra 'hybrid' object (every rank's shard of configs[4] through the hybrid pipeline)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    from legal_rag_b200 import engine

    peaks = measured_peaks()
    wl = WORKLOADS[args.workload](args, rank, world, device)
    line = measure_native(wl, args.steps, args.warmup, rank, world, local_rank, device, peaks)

    # The metric's own multi-GPU configuration next to the configs[1] line: every rank's 1/8 shard of configs[4] through the
    # sharded hybrid pipeline (a shorter timed region; same rules).  8 ranks hold exactly configs[4].
    if args.workload == "dense" and not args.no_hybrid_block and not (args.n_docs or args.dim or args.nq):
        wl_h = WORKLOADS["hybrid"](args, rank, world, device)
        release(wl)
        try:
            h = measure_native(wl_h, max(1, min(args.steps, 5)), args.warmup, rank, world, local_rank, device, peaks)
        except Exception as exc:       # the configs[1] line above stands on its own
            h = {"error": f"{type(exc).__name__}: {exc}"}
        release(wl_h)
        if rank == 0:
            line["hybrid"] = h if "error" in h else {key: h[key] for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config",
                                                      "clocks", "e2e", "gpu_launches", "roofline", "global_queries_per_s")}

    if rank == 0:
        if not args.no_cpu_baseline:
            run, sample = wl.cpu_sample()
            run()
            vals, t0 = [], time.perf_counter()
            while time.perf_counter() - t0 < 12.0 and len(vals) < 20:
                vals.append(run()[0])
            v = len(vals) / sum(1.0 / x for x in vals)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": getattr(wl, "cpu_cores", os.cpu_count()), "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
