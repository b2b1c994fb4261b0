#!/usr/bin/env python
"""Benchmark of the hybrid-retrieval hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hybrid|dense|maxsim|maxsim_scan|bm25|ucc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic queries.  The default workload is the metric's own:
BASELINE.json configs[4], the full hybrid step (dense 768-d flat-IP + BM25 + ColBERT MaxSim rerank + weighted fusion, k = 100,
4096 queries).  configs[4] is 100M documents over 8 GPUs and does not fit one: every rank owns 1/8 of it (12.5M docs: weak
scaling, 8 ranks hold exactly configs[4]), computes local top-k lists, and the lists meet in one NCCL all-gather + merge per
step.  The JSON line (rank 0) carries, besides the contract's keys:
  roofline        the LONGEST stage of the step against its bound, and `stages`: every kernel family of the step with its own
                  algorithmic work, device time (CUDA events on the launching stream) and fraction of the measured peak
  parity_check    32 sampled queries recomputed with plain torch on every rank and merged (bench_parity.py): ids at N GPUs
  strong          (N > 1) the same per-GPU shard split over the N ranks: ms/step, speed-up, what the rest of the step costs
  kernels         (N = 1) short runs of the other configs (configs[1] dense, configs[2] MaxSim, configs[3] BM25, full-scan MaxSim,
                  top-k select shapes), each with its roofline, so that every kernel family has a driver-run number
  api             (N = 1) the reference's class API on configs[0]'s corpus: HybridRetriever.search / search_batch
  cpu_baseline    the oracle port of the reference's path on the host cores, bounded sample
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
            entries = json.load(f).get(kernel) or []
        for entry in entries if isinstance(entries, list) else [entries]:
            if all(entry["shape"].get(k) == v for k, v in shape.items()):
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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sm_mhz_min": min(sm) if sm else None,
                "sm_mhz_max_seen": max(sm) if sm else None, "power_w": statistics.median(pw) if pw else None,
                "power_w_max": max(pw) if pw else None}


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
        scan_gbs = self.N * self.d * 2 / (kernel_ms * 1e-3) / 1e9
        if self.nq < 216:
            # below the ridge point (216 FLOP per corpus byte) the scan is bound by reading the corpus once
            return {"bound": "hbm", "achieved": scan_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": scan_gbs / peaks["hbm_gbs"],
                    "traffic": None, "kernel": "dense_scan_kernel", "kernel_ms": kernel_ms,
                    "algorithmic": f"2*N*d = {2.0 * self.N * self.d:.3e} B per step (the corpus, once)", "peak_source": peaks["source"] + " (copy bandwidth)",
                    "scan_gbs": scan_gbs}
        ach = flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "traffic": measured_traffic("dense_scan_kernel", N=self.N, d=self.d, nq=self.nq, k=self.k),
                "kernel": "dense_scan_kernel", "kernel_ms": kernel_ms,
                "algorithmic": f"2*nq*N*d = {flops:.3e} FLOP per launch", "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                "scan_gbs": scan_gbs}

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

    def __init__(self, args, rank, world, device, n_docs=None):
        self.N, self.d, self.V = n_docs or args.n_docs or 12_500_000, args.dim or 768, args.vocab or 500_000
        self.nq, self.k, self.kc = args.nq or 4096, args.k, (getattr(args, "kc", 0) or args.k)
        self.colbert_mode = getattr(args, "colbert_mode", "rerank")
        self.dense_sms_arg = str(getattr(args, "dense_sms", "auto"))
        self.partition, self.alone_ms = None, {}
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
                "parallelism": (f"doc-sharded x{self.world}: per channel local top-k, ONE packed NCCL all-gather, merge kernels that read the "
                                f"gathered buffers in place, MaxSim by row owner + NCCL max-reduce, replicated fusion") if self.world > 1 else "single GPU",
                "l2": f"dense shard {self.N * self.d * 2 / 1e9:.1f} GB, postings {getattr(self, 'nnz', 0) * 8 / 1e9:.1f} GB >> 126 MB L2 (no flush needed)",
                "stages_ms": getattr(self, "stages_ms", None),
                "sm_partition": self.partition,
                "value_definition": "weak scaling: every rank answers the whole batch against its own 1/8 shard of configs[4]; "
                                    "value = n_gpus * batch_queries * steps / time (shard-level query scans per second); "
                                    "global_queries_per_s = batch_queries * steps / time is the rate at which the n_gpus-shard corpus answers queries"}

    def setup(self):
        import numpy as np
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
        # exact algorithmic posting bytes of this shard's BM25 stage: 8 B x sum over queries of df over the query's distinct terms
        df = st["df"].cpu().numpy()
        qi, qt = self.q_indptr.cpu().numpy(), self.q_term.cpu().numpy()
        self.alg_postings = int(sum(int(df[np.unique(qt[qi[j]:qi[j + 1]])].sum()) for j in range(self.nq)))
        self.cand_owned = None
        self._stage_times()
        self._pick_partition()

    def _stage_times(self):
        """Per-stage device times of one step with the stages ONE AFTER THE OTHER, each on the whole machine, run stage by stage
        with a synchronising event after each (outside the timed region), reported in config.stages_ms; the library's
        launch profiler gives every kernel family's time in that arrangement (`alone_ms`).  Also counts the MaxSim candidates
        this rank scores."""
        torch, eng, sh = self.torch, self.engine, self.shard
        ev = lambda: torch.cuda.Event(enable_timing=True)
        sh.dense_sms = 0
        for it in range(2):
            if it == 1:
                eng.prof_enable(64)
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
                self.cand_owned = int(own.sum())
                eng.maxsim_scores(sh.tokens, None, self.Qtok, torch.where(own, rows, torch.full_like(rows, -1)))
            else:
                eng.allgather_merge(*eng.maxsim_scan_topk(sh.tokens, None, self.Qtok, self.kc, id_base=sh.id_base), self.kc)
            marks[3].record()
            if it == 1:
                torch.cuda.synchronize()
                self.alone_ms = {}
                for name, t in eng.prof_collect():
                    self.alone_ms[name] = self.alone_ms.get(name, 0.0) + t
                eng.prof_enable(0)
            self.step()
            marks[4].record()
            torch.cuda.synchronize()
        self.stages_ms = {"dense+merge": marks[0].elapsed_time(marks[1]), "bm25+merge": marks[1].elapsed_time(marks[2]),
                          "candidates+maxsim": marks[2].elapsed_time(marks[3]), "whole_step": marks[3].elapsed_time(marks[4]),
                          "arrangement": "one after the other, each stage on the whole GPU"}

    def _pick_partition(self):
        """SM partition of the two big scans (HybridShard.dense_sms, csrc/partition.cu): `--dense-sms N` fixes it, `0` runs the
        stages one after the other, `auto` (default) times a few splits for three steps each, outside the timed region, and
        keeps the fastest (max over ranks)."""
        torch, sh = self.torch, self.shard
        import torch.distributed as dist
        sms = self.engine._native.load().lrag_sm_count()
        if self.dense_sms_arg != "auto":
            sh.dense_sms = max(0, min(int(self.dense_sms_arg), sms - 8))
            tried = None
        else:
            timed = sh.tune_partition(self.Qd, self.q_indptr, self.q_term, self.mx, self.Qtok, k=self.k, kc=self.kc,
                                      colbert_mode=self.colbert_mode)
            tried = {("one after the other" if c == 0 else f"dense on {c} SMs"): round(ms, 2) for c, ms in timed.items()}
        D = sh.dense_sms
        self.partition = {"dense_sms": D, "bm25_sms": sms - D if D else 0, "sms": sms, "tried_ms_per_step": tried,
                          "what": ("dense scan and BM25 scan side by side on disjoint SM sets (lrag_sm_reserve): the power-capped dense stage "
                                   "no longer leaves the low-power BM25 stage a throttled clock; results are identical to the serial step")
                                  if D else "stages one after the other, each on the whole GPU"}

    def step(self):
        return self.shard.search_device(self.Qd, self.q_indptr, self.q_term, self.mx, self.Qtok, k=self.k, kc=self.kc,
                                        colbert_mode=self.colbert_mode)

    def e2e_step(self):
        return self.shard.search(self.host[0], self.host[1], self.host[2], self.mx, self.host[3], k=self.k, kc=self.kc,
                                 colbert_mode=self.colbert_mode)

    def critical_kernel_ms(self, by_tag):
        """Kernel time on the step's critical path: side-by-side scans count once (the longer one)."""
        if self.partition and self.partition.get("dense_sms"):
            both = ("dense_scan", "bm25_scan")
            return max(by_tag.get(t, 0.0) for t in both) + sum(v for t, v in by_tag.items() if t not in both)
        return sum(by_tag.values())

    def e2e_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.host), self.nq * self.k * 12

    def units_per_step(self):
        return self.nq * self.world

    def stage_rooflines(self, by_tag, peaks):
        """One entry per kernel family of the step: algorithmic work per step, device time per step (summed over the family's
        launches), achieved rate and fraction of the measured peak that bounds it."""
        out = []
        part = self.partition or {}
        share = {"dense_scan": part.get("dense_sms", 0), "bm25_scan": part.get("bm25_sms", 0)} if part.get("dense_sms") else {}
        def entry(kernel, tag, bound, work, unit_div, unit, peak, alg, **extra):
            ms = by_tag.get(tag, 0.0)
            if ms <= 0:
                return
            ach = work / (ms * 1e-3) / unit_div
            e = dict({"kernel": kernel, "bound": bound, "kernel_ms": ms, "achieved": ach, "peak": peak, "unit": unit,
                      "frac": ach / peak, "algorithmic": alg}, **extra)
            if share.get(tag):
                # the kernel ran on a part of the machine, side by side with the other scan: `frac` is still against the
                # whole GPU's peak; per SM it is frac / sm_share; `alone` = the same kernel by itself on the whole GPU,
                # measured in this run with the stages one after the other (setup, outside the timed region)
                e["sms"], e["sm_share"] = share[tag], share[tag] / part["sms"]
                e["frac_of_sm_share"] = e["frac"] / e["sm_share"]
                e["note"] = (f"the kernel runs on {share[tag]} of {part['sms']} SMs while the other scan has the rest: `frac` divides by the WHOLE "
                             f"GPU's peak, `frac_of_sm_share` by the kernel's share of it, `alone` is the same kernel by itself on the whole "
                             f"GPU in this run's serial arrangement (after the power-capped dense stage)")
                if self.alone_ms.get(tag):
                    a = work / (self.alone_ms[tag] * 1e-3) / unit_div
                    e["alone"] = {"kernel_ms": self.alone_ms[tag], "achieved": a, "frac": a / peak}
            out.append(e)
        flops = 2.0 * self.nq * self.N * self.d
        entry("dense_scan_kernel", "dense_scan", "tensor", flops, 1e12, "TFLOP/s", peaks["bf16_tflops"], f"2*nq*N*d = {flops:.3e} FLOP",
              peak_burst=peaks["bf16_tflops_burst"])
        pbytes = 8.0 * self.alg_postings
        entry("bm25_scan_kernel", "bm25_scan", "hbm", pbytes, 1e9, "GB/s", peaks["hbm_gbs"],
              f"8 B x sum_q sum_(distinct t in q) df(t) = {pbytes:.4e} B (served from L2 after the first touch: a posting-throughput "
              f"figure quoted against the HBM copy peak, see DESIGN 4.2)",
              traffic=measured_traffic("bm25_scan_kernel", N=self.N, V=self.V, nq=self.nq, k=self.kc, mean_len=24.0))
        if self.colbert_mode == "rerank" and self.cand_owned:
            mbytes = float(self.cand_owned) * self.Ld * 128 * 2
            entry("maxsim_kernel", "maxsim", "hbm", mbytes, 1e9, "GB/s", peaks["hbm_gbs"],
                  f"owned candidates x Ld x dim x 2 = {self.cand_owned} x {self.Ld} x 128 x 2 = {mbytes:.3e} B")
        else:
            sflops = 2.0 * self.nq * self.Lq * 128 * self.Nd_tok * self.Ld
            entry("maxsim_scan_kernel", "maxsim_scan", "tensor", sflops, 1e12, "TFLOP/s", peaks["bf16_tflops"], f"2*nq*Lq*dim*Nd*Ld = {sflops:.3e} FLOP")
        fbytes = float(self.nq) * (2 * self.kc * 12 + 2 * self.kc * 12) + float(self.nq) * (3 * self.kc * 12 + self.k * 12)
        entry("fuse_kernel", "fuse", "hbm", fbytes, 1e9, "GB/s", peaks["hbm_gbs"],
              f"two launches: candidate union (2 lists in, 2 kc out) + final fusion (3 lists in, k out) = {fbytes:.3e} B; latency-bound at "
              f"this size (DESIGN 4.4)")
        sbytes = float(self.nq) * 2 * self.kc * (4 + 8 + 12)
        entry("topk_select_kernel", "select", "hbm", sbytes, 1e9, "GB/s", peaks["hbm_gbs"],
              f"MaxSim scores + candidate ids in, kc hits out = {sbytes:.3e} B; latency-bound at this size")
        return out

    def roofline(self, by_tag, peaks):
        stages = self.stage_rooflines(by_tag, peaks)
        top = max(stages, key=lambda e: e["kernel_ms"])
        roof = dict(top)
        roof.setdefault("traffic", None)
        roof["kernel"] = top["kernel"] + (f" on {top['sms']} of {self.partition['sms']} SMs, side by side with the other scan (longest kernel of "
                                          f"the hybrid step)" if top.get("sms") else " (longest stage of the hybrid step)")
        roof["peak_source"] = peaks["source"] + (" (copy bandwidth)" if top["bound"] == "hbm" else " (sustained cuBLAS bf16)")
        roof["stages"] = stages
        return roof

    # ---- self-check: 32 sampled queries through plain torch on every rank (bench_parity.py) ----
    def parity_check(self, n_sample=32):
        import bench_parity as bp
        torch, eng, sh = self.torch, self.engine, self.shard
        rows = torch.linspace(0, self.nq - 1, n_sample, device=self.device).long().unique()
        rl = rows.tolist()
        kc, k, K = self.kc, self.k, self.kc + 8
        # the engine's per-channel lists and final result for the whole batch (all ranks take part in the collectives)
        ed = eng.allgather_merge(*eng.dense_topk(sh.X, self.Qd, kc, sh.id_base), kc)
        eb = eng.allgather_merge(*eng.bm25_topk(sh.bm25, self.q_indptr, self.q_term, self.mx, kc), kc)
        es, ei = self.step()
        # the checker's
        rd = bp.dense_lists(sh.X, self.Qd[rows], K, sh.id_base)
        rb = bp.bm25_lists(sh.bm25, self.q_indptr, self.q_term, rl, K)
        out = {"queries": len(rl), "k": k, "checker": "torch fp32 matmul / fp64 scatter-add / fp32 einsum / fp64 min-max fusion per rank, "
                                                     "merged with torch all_gather + sorts (bench_parity.py)"}
        chans = {}
        for name, (gs, gi), (rs, ri), tau in (("dense", ed, rd, 1e-3), ("bm25", eb, rb, 1e-3)):
            dec, mis, rel = bp.compare(gs[rows], gi[rows], rs, ri, tau)
            chans[name] = {"decided": dec, "mismatched_ids": mis, "max_rel_err": rel, "tau": tau}
        # the rest of the pipeline from the checker's own channel lists
        rdl, rbl = (rd[0][:, :kc], rd[1][:, :kc]), (rb[0][:, :kc], rb[1][:, :kc])
        lists = [rdl, rbl]
        weights = [0.6, 0.4]
        if self.colbert_mode == "rerank":
            _, cand = bp.minmax_weighted_sum([rdl, rbl], [0.6, 0.4], 2 * kc)
            rc = bp.maxsim_candidates(sh.tokens, self.Qtok[rows], cand, sh.tok_row_base, sh.tok_rows_total, kc)
            lists.append(rc)
            weights.append(0.35)
        rf = bp.minmax_weighted_sum(lists, weights, k + 8)
        dec, mis, rel = bp.compare(es[rows], ei[rows], rf[0], rf[1], 1e-3)
        chans["fused"] = {"decided": dec, "mismatched_ids": mis, "max_rel_err": rel, "tau": 1e-3}
        out["channels"] = chans
        out["mismatched_ids"] = sum(c["mismatched_ids"] for c in chans.values())
        out["max_rel_err"] = max(c["max_rel_err"] for c in chans.values())
        return out

    def cpu_sample(self, budget_s=15.0):
        """The reference's sequential path (hybrid_retriever.py:282-384) restated: flat-IP + literal BM25 + MaxSim of the
        candidate union + _fuse, one query at a time per worker process, on a 20k-doc sample; dense and BM25 extrapolated
        linearly in docs.  The reference itself is single-process; one worker per host core is the most the host can do."""
        import numpy as np
        from oracle import bm25 as obm25, dense as odense, fuse as ofuse, maxsim as omaxsim
        n_s, v_s = 20_000, 5_000
        workers = max(1, os.cpu_count() or 1)
        nq_s = 2 * workers
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
        kc, N = self.kc, self.N

        def one(j):
            t0 = time.perf_counter()
            ds, di = odense.flat_ip_topk(Q[j:j + 1], X, kc)
            bs, bi = obm25.search(lit, queries[j], kc)
            t1 = time.perf_counter()
            dl = list(zip(di[0].tolist(), ds[0].tolist())); bl = list(zip(bi.tolist(), bs.tolist()))
            cand = np.array([[r["id"] % 2000 for r in ofuse.fuse(dl, bl, [], method="weighted_sum")]])
            cs = omaxsim.maxsim_scores(Qt[j:j + 1], D, None, cand)[0]
            order = np.argsort(-cs, kind="stable")[:kc]
            ofuse.fuse(dl, bl, [(int(cand[0][o]), float(cs[o])) for o in order], method="weighted_sum")
            t2 = time.perf_counter()
            return t1 - t0, t2 - t1
        global _CPU_ONE
        _CPU_ONE = one

        def run():
            import multiprocessing as mp
            t0 = time.perf_counter()
            if workers > 1:
                with mp.get_context("fork").Pool(workers) as pool:         # forked workers share the sample corpus
                    parts = pool.map(_cpu_one, range(nq_s))
            else:
                parts = [one(j) for j in range(nq_s)]
            wall = time.perf_counter() - t0
            # per-query cost on the full shard: the scans scale linearly in documents, the rest does not
            t_scan, t_rest = sum(a for a, _ in parts), sum(b for _, b in parts)
            per_query_full = (t_scan * N / n_s + t_rest) / nq_s
            return workers / per_query_full, wall
        return run, (f"oracle restatement of the reference's sequential hybrid path, one query at a time in each of {workers} worker processes: "
                     f"numpy flat-IP + literal BM25Okapi over {n_s} docs (extrapolated linearly to {N}), MaxSim of the candidate union, _fuse; "
                     f"{nq_s} queries per step")

    @property
    def cpu_cores(self):
        return max(1, os.cpu_count() or 1)


_CPU_ONE = None


def _cpu_one(j):
    return _CPU_ONE(j)


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
        self.ragged = bool(getattr(args, "ragged_doclen", False))
        self.doclen = None

    def config(self):
        return {"workload": f"configs[2] token store, full scan: {self.Nd} docs x {self.Ld} x 128 bf16, batch of {self.nq} queries x {self.Lq} tokens "
                            f"against every document -> top-{self.k}" + (", document lengths ~U{32..128} (masked epilogue; FLOPs counted on the "
                                                                        "padded store, which is what the tensor cores multiply)" if self.ragged else ""),
                "l2": f"token store {self.Nd * self.Ld * 256 / 1e9:.1f} GB >> L2",
                "parallelism": f"doc-sharded x{self.world}" if self.world > 1 else "single GPU"}

    def setup(self):
        import torch
        from legal_rag_b200 import engine, synth
        self.torch, self.engine = torch, engine
        self.D = synth.unit_tokens_bf16(self.Nd, self.Ld, 128, 5 + self.rank, self.device)
        self.Q = synth.unit_tokens_bf16(self.nq, self.Lq, 128, 7, self.device)
        self.Q_host = self.Q.cpu().pin_memory()
        if self.ragged:                          # SURVEY 8d C3's masking variant: doclen ~U{32..128}, seed 6
            g = torch.Generator(device=self.device).manual_seed(6 + self.rank)
            self.doclen = torch.randint(32, self.Ld + 1, (self.Nd,), generator=g, device=self.device, dtype=torch.int32)

    def step(self):
        s, i = self.engine.maxsim_scan_topk(self.D, self.doclen, self.Q, self.k, id_base=self.rank * self.Nd)
        return self.engine.allgather_merge(s, i, self.k)

    def e2e_step(self):
        q = self.Q_host.to(self.device, non_blocking=True)
        s, i = self.engine.maxsim_scan_topk(self.D, self.doclen, q, self.k, id_base=self.rank * self.Nd)
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
def pin_host_threads():
    """The reference arm runs under torchrun at N > 1, which exports OMP_NUM_THREADS=1: numpy's BLAS would run on one thread
    while the line claims all cores (round-1 finding).  Set the thread counts explicitly before numpy is imported and report
    what the BLAS pool actually uses."""
    n = str(max(1, os.cpu_count() or 1))
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[var] = n
    import numpy  # noqa: F401
    used = int(n)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=int(n))
        pools = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        if pools:
            used = max(pools)
    except Exception:
        pass
    return used


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port; the third-party wheels it calls are not
    installable here) on the host cores."""
    if rank != 0:
        return
    blas_threads = pin_host_threads()

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
    # "shard-level query scans per second" does not depend on N for the CPU path.
    wl, value, ms, sample = timed(args.workload)
    cores = getattr(wl, "cpu_cores", blas_threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.config(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "blas_threads": blas_threads, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def with_burst(roof, peaks):
    """Tensor-bound rooflines are quoted against the sustained cuBLAS figure (the kernel runs inside a long, power-capped
    step); the same achieved rate against the burst figure is reported next to it -- a stage that follows a cooler stage
    (the dense scan inside the hybrid step) runs at clocks between the two and can exceed the sustained number."""
    if roof.get("bound") == "tensor":
        roof["peak_burst"] = peaks["bf16_tflops_burst"]
        roof["frac_of_burst"] = roof["achieved"] / peaks["bf16_tflops_burst"]
    for st in roof.get("stages", []):
        if st.get("bound") == "tensor":
            st["peak_burst"] = peaks["bf16_tflops_burst"]
            st["frac_of_burst"] = st["achieved"] / peaks["bf16_tflops_burst"]
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


def measure_native(wl, steps, warmup, rank, world, local_rank, device, peaks, e2e=True, clocks=True):
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
    engine.prof_enable(steps * 24 + 8)
    launches0 = engine.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0 and clocks:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        wl.step()
    ev1.record()
    barrier()
    clk = sampler.stop() if rank == 0 and clocks else None
    ms = ev0.elapsed_time(ev1)
    launches = engine.launch_count() - launches0
    prof = engine.prof_collect()
    engine.prof_enable(0)
    by_tag = {}
    for name, t in prof:
        by_tag[name] = by_tag.get(name, 0.0) + t / steps      # per step (a step may launch a kernel family several times)
    tags = sorted(by_tag)

    # ---- end to end through the host-buffer API: H2D queries -> search -> D2H results, every step ----
    e2e_ms = 0.0
    if e2e:
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

    t = torch.tensor([ms, e2e_ms] + [by_tag[n] for n in tags], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vals = t.tolist()
    ms, e2e_ms = vals[0], vals[1]
    by_tag = dict(zip(tags, vals[2:]))
    if rank != 0:
        return None
    units = wl.units_per_step() * steps
    roof = wl.roofline(by_tag, peaks) if hasattr(wl, "stage_rooflines") else wl.roofline(by_tag.get(wl.dominant, 0.0), peaks)
    line = {"metric": METRIC, "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic (seeded, generated on device; random unit-norm vectors, Zipfian postings)", "config": wl.config(),
            "clocks": clk, "gpu_launches": launches,
            "roofline": with_burst(roof, peaks),
            "global_queries_per_s": wl.nq * steps / (ms * 1e-3),
            "kernel_ms_per_step": by_tag,
            "rest_of_step_ms": ms / steps - (wl.critical_kernel_ms(by_tag) if hasattr(wl, "critical_kernel_ms") else sum(by_tag.values()))}
    if roof.get("stages"):
        # the step as a whole against its stages' rooflines: the time a machine that ran every stage alone at the measured
        # peak bounding it would need for one step, over the measured step.  With the two scans side by side neither kernel's
        # own `frac` (against the WHOLE GPU's peak) can approach 1 -- they share the machine -- but their sum can.
        ideal = {st["kernel"]: st["kernel_ms"] * st["frac"] for st in roof["stages"]}
        roof["step"] = {"ideal_ms": sum(ideal.values()), "ideal_ms_by_stage": ideal, "ms_per_step": ms / steps,
                        "frac": sum(ideal.values()) / (ms / steps),
                        "what": "sum over stages of (algorithmic work / the measured peak that bounds the stage) / measured step time"}
    if e2e:
        h2d, d2h = wl.e2e_bytes()
        line["e2e"] = {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_ms / steps}
    return line


# =================================================================================================
# side blocks of the default line
# =================================================================================================
def kernels_block(args, device, peaks):
    """Short runs of the other configs on one GPU, so that every kernel family has a driver-run roofline number."""
    import copy
    import torch
    from legal_rag_b200 import engine
    out = {}

    def one(name, steps, **over):
        a = copy.copy(args)
        a.n_docs = a.dim = a.nq = a.vocab = 0
        a.mean_len = 0.0
        for key, v in over.items():
            setattr(a, key, v)
        wl = WORKLOADS[name](a, 0, 1, device)
        try:
            line = measure_native(wl, steps, 3, 0, 1, 0, device, peaks, e2e=False, clocks=False)
            r = line["roofline"]
            out[wl.name] = {"workload": line["config"]["workload"], "ms_per_step": line["ms_per_step"], "queries_per_s": line["value"],
                            "roofline": {k: r[k] for k in ("kernel", "bound", "kernel_ms", "achieved", "peak", "unit", "frac", "algorithmic", "traffic")
                                         if k in r}}
            if "frac_of_burst" in r:
                out[wl.name]["roofline"]["frac_of_burst"] = r["frac_of_burst"]
            if "scan_gbs" in r:
                out[wl.name]["roofline"]["scan_gbs"] = r["scan_gbs"]
        except Exception as exc:
            out[wl.name] = {"error": f"{type(exc).__name__}: {exc}"}
        release(wl)

    one("dense", 5)                       # configs[1]
    cfg1 = out.pop("dense_flat_ip_top100", None)
    one("dense", 5, nq=64)                # the same corpus, HBM-bound batch: scan GB/s
    out["dense_flat_ip_top100_nq64"] = out.pop("dense_flat_ip_top100", None)
    out["dense_flat_ip_top100"] = cfg1
    one("maxsim", 5)                      # configs[2]
    one("maxsim_scan", 3)                 # configs[2]'s store, every document
    one("bm25", 3)                        # configs[3]
    # top-k select over materialised score rows: the shapes the pipeline produces
    sel = {}
    for nq, n in ((256, 1 << 20), (64, 1 << 20), (4096, 20_000)):
        S = torch.randn((nq, n), device=device)
        for _ in range(3):
            engine.topk_select(S, 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            engine.topk_select(S, 100)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = 4.0 * nq * n / (ms * 1e-3) / 1e9
        sel[f"{nq}x{n}"] = {"ms": ms, "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "bound": "hbm",
                            "algorithmic": f"4*nq*N = {4.0 * nq * n:.3e} B (select + merge launches, CUDA events around both)"}
        del S
    out["topk_select_k100"] = sel
    return out


def strong_block(args, rank, world, local_rank, device, peaks, weak_ms):
    """Strong scaling of the hybrid step: the per-GPU shard of the weak run (what ONE GPU holds) split over the N ranks."""
    import copy
    a = copy.copy(args)
    per = (12_500_000 if not args.n_docs else args.n_docs) // world
    wl = HybridWorkload(a, rank, world, device, n_docs=per)
    line = measure_native(wl, max(2, min(args.steps, 5)), 3, rank, world, local_rank, device, peaks, e2e=False, clocks=False)
    release(wl)
    if rank != 0:
        return None
    ms = line["ms_per_step"]
    rest = line["rest_of_step_ms"]
    return {"corpus_docs_total": per * world, "docs_per_gpu": per, "ms_per_step": ms, "queries_per_s": line["global_queries_per_s"],
            "one_gpu_ms_per_step": weak_ms, "speedup_vs_one_gpu": weak_ms / ms, "efficiency": weak_ms / ms / world,
            "kernel_ms_per_step": line["kernel_ms_per_step"], "collectives_and_glue_ms": rest,
            "limiter": ("kernel time falls with 1/N; what does not shrink is the exchange: one packed all-gather + one max-reduce per step, "
                        "the merge / fusion / select launches over nq x k lists and their launch gaps (collectives_and_glue_ms)"),
            "note": "one_gpu_ms_per_step is this run's weak-scaling step (every rank scans a full single-GPU shard)"}


def api_block(args, device):
    """The reference's class API on configs[0]'s corpus (UCC articles; hashing stand-ins for the BGE / jieba encoders, whose
    forward passes are not part of the path): HybridRetriever.search(question, None, 100) one query at a time and
    HybridRetriever.search_batch, through artifacts written by this package's builders."""
    import tempfile
    import numpy as np
    from legal_rag_b200.config import AppConfig
    from legal_rag_b200.retrieval import HybridRetriever, builders, encoders
    from legal_rag_b200.schemas import LawChunk
    z = np.load(os.path.join(ROOT, "tests", "golden", "ucc_corpus.npz"))
    lens, flat, vocab, ids = z["doc_len"], z["tokens"].astype(np.int64), z["vocab"], z["ids"]
    off = np.concatenate([[0], np.cumsum(lens)])
    texts = [" ".join(vocab[flat[off[i]:off[i + 1]]]) for i in range(len(lens))]
    chunks = [LawChunk(id=str(ids[i]), law_name="UCC", article_no=str(i), article_id=str(i), text=texts[i], lang="en")
              for i in range(len(texts))]
    rng = np.random.default_rng(44)
    questions = []
    for _ in range(256):
        t = texts[int(rng.integers(0, len(texts)))].split()
        L = int(rng.integers(3, 9))
        st = int(rng.integers(0, max(1, len(t) - L)))
        questions.append(" ".join(t[st:st + L]))
    out = {"corpus": f"configs[0]: {len(chunks)} UCC articles", "encoders": "HashingDenseEncoder(768) / regex tokenizer (stand-ins)"}
    with tempfile.TemporaryDirectory() as root:
        cfg = AppConfig()
        cfg.device = str(device)
        r = cfg.retrieval
        r.faiss_index_file, r.faiss_meta_file = os.path.join(root, "faiss", "faiss.index"), os.path.join(root, "faiss", "faiss_meta.jsonl")
        r.bm25_index_file = os.path.join(root, "bm25.pkl")
        r.enable_colbert = False
        r.top_k, r.min_final_score, r.fusion_method = 100, 0.0, "weighted_sum"
        enc = encoders.HashingDenseEncoder(768)
        encoders.register_dense_encoder(lambda name, dev: enc)
        try:
            builders.build_faiss_index(cfg, chunks)
            builders.build_bm25_index(cfg, chunks)
            hr = HybridRetriever(cfg)
            for q in questions[:8]:
                hr.search(q, None, 100)
            t0 = time.perf_counter()
            for q in questions[:64]:
                hits = hr.search(q, None, 100)
            dt = time.perf_counter() - t0
            out["search_us_per_query"] = 1e6 * dt / 64
            out["search_hits"] = len(hits)
            hr.search_batch(questions, 100)
            t0 = time.perf_counter()
            for _ in range(3):
                res = hr.search_batch(questions, 100)
            dt = time.perf_counter() - t0
            out["search_batch_queries_per_s"] = 3 * len(questions) / dt
            out["search_batch_size"] = len(questions)
            out["fast_path"] = bool(getattr(hr, "fast_path_used", False))
        finally:
            encoders.register_dense_encoder(None)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="hybrid", choices=sorted(WORKLOADS))
    ap.add_argument("--n-docs", type=int, default=0)
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--vocab", type=int, default=0)
    ap.add_argument("--nq", type=int, default=0)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--mean-len", type=float, default=0.0, help="bm25: mean document length of the synthetic corpus (default 40)")
    ap.add_argument("--kc", type=int, default=0, help="hybrid: per-channel list length (default k); 500 gives the 1000-candidate rerank of SURVEY 8d C5")
    ap.add_argument("--colbert-mode", default="rerank", choices=["rerank", "scan"],
                    help="hybrid: MaxSim over the fused candidate union, or ColBERT as a first-stage channel over the whole token store")
    ap.add_argument("--ragged-doclen", action="store_true", help="maxsim_scan: document lengths ~U{32..128} instead of full-length documents")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-sms", default="auto", help="hybrid workload: SMs of the dense scan when it runs side by side with the BM25 "
                    "scan (the rest go to BM25); 0 = one after the other; auto = time a few splits and keep the fastest")
    ap.add_argument("--no-side-blocks", action="store_true",
                    help="default workload only: skip parity_check / strong / kernels / api (profiling runs)")
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

    peaks = measured_peaks()
    wl = WORKLOADS[args.workload](args, rank, world, device)
    line = measure_native(wl, args.steps, args.warmup, rank, world, local_rank, device, peaks)
    default_shape = args.workload == "hybrid" and not (args.n_docs or args.dim or args.nq or args.vocab) and not args.no_side_blocks

    def guarded(name, fn):
        try:
            return fn()
        except Exception as exc:       # the main line stands on its own
            import traceback
            traceback.print_exc()
            return {"error": f"{type(exc).__name__}: {exc}"}

    if hasattr(wl, "parity_check") and not args.no_side_blocks:
        pc = guarded("parity_check", wl.parity_check)           # every rank takes part in its collectives
        if rank == 0:
            line["parity_check"] = pc
    release(wl)
    if default_shape and world > 1:
        sb = guarded("strong", lambda: strong_block(args, rank, world, local_rank, device, peaks, line["ms_per_step"] if rank == 0 else 0.0))
        if rank == 0:
            line["strong"] = sb
    if default_shape and world == 1:
        line["kernels"] = guarded("kernels", lambda: kernels_block(args, device, peaks))
        line["api"] = guarded("api", lambda: api_block(args, device))

    if rank == 0:
        if not args.no_cpu_baseline:
            blas_threads = pin_host_threads()
            run, sample = wl.cpu_sample()
            run()
            vals, t0 = [], time.perf_counter()
            while time.perf_counter() - t0 < 12.0 and len(vals) < 20:
                vals.append(run()[0])
            v = len(vals) / sum(1.0 / x for x in vals)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": getattr(wl, "cpu_cores", blas_threads), "kind": "port", "sample": sample,
                                    "blas_threads": blas_threads}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
