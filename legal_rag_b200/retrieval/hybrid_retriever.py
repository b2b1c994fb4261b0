"""HybridRetriever (replaces legalrag/retrieval/hybrid_retriever.py:133-551): same channels, same four
fusion methods, same score_breakdown, with every scoring step on the GPU.  `search(question, llm, top_k,
decision)` keeps the reference's positional order; `search_batch` is the throughput entry point (all
queries through every channel in one kernel launch each, fused on the device)."""
from __future__ import annotations

import threading

import logging
import time
import traceback
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Set

import numpy as np
import torch

from .. import engine
from ..schemas import RetrievalHit
from .bm25_retriever import BM25Retriever
from .colbert_retriever import ColBERTRetriever
from .dense_retriever import DenseRetriever

logger = logging.getLogger("legal_rag_b200.retrieval")
CHANNELS = ("dense", "bm25", "colbert")


def _as_channel_list(x: Any) -> List[str]:
    """score_breakdown["channel"] in any of the shapes the reference tolerates -> list of names."""
    if x is None:
        return []
    return [str(v) for v in x] if isinstance(x, (list, set, tuple)) else [str(x)]


def _by_score(hits: List[RetrievalHit]) -> List[RetrievalHit]:
    """In-place: best score first (stable), ranks 1..n."""
    hits.sort(key=lambda h: -float(h.score))
    for pos, h in enumerate(hits):
        h.rank = pos + 1
    return hits


def _dedup_keep_best(hits: List[RetrievalHit]) -> List[RetrievalHit]:
    """One hit per chunk id (hybrid_retriever.py:71-130 of the reference, checked against its output in
    tests/golden/fuse_golden.json): the occurrence with the highest score represents the chunk (the earliest one on equal
    scores) and inherits the provenance of all of them -- the union of their `channel` lists and the per-channel sum of
    their `channel_contrib` dicts; channels are listed by summed contribution (by name when there is none).  Chunks come
    out best score first, in first-appearance order on ties."""
    occurrences: Dict[str, List[RetrievalHit]] = {}
    for h in hits:
        occurrences.setdefault(h.chunk.id, []).append(h)
    kept: List[RetrievalHit] = []
    for group in occurrences.values():
        rep = group[0]
        for h in group[1:]:
            if float(h.score) > float(rep.score):
                rep = h
        info = rep.score_breakdown or {}
        if len(group) == 1:
            if "channel" in info:
                info["channel"] = _as_channel_list(info["channel"])
                rep.score_breakdown = info
        else:
            names: Set[str] = set()
            share: Dict[str, float] = {}
            for h in group:
                sb = h.score_breakdown or {}
                names.update(_as_channel_list(sb.get("channel")))
                part = sb.get("channel_contrib") or {}
                if isinstance(part, dict):
                    for name, value in part.items():
                        share[str(name)] = share.get(str(name), 0.0) + float(value)
            if share:
                info["channel_contrib"] = share
            info["channel"] = sorted(names, key=lambda c: (-share.get(c, 0.0), c)) if share else sorted(names)
            rep.score_breakdown = info
        kept.append(rep)
    return _by_score(kept)


class _Laps:
    """Wall-clock laps of one search() for the reference's timing log line."""

    def __init__(self) -> None:
        self.t = {"start": time.time()}

    def mark(self, name: str) -> None:
        self.t[name] = time.time()

    def ms(self, a: str, b: str) -> int:
        return int((self.t[b] - self.t[a]) * 1000) if a in self.t and b in self.t else 0


@dataclass
class HybridRetriever:
    cfg: Any

    def __post_init__(self) -> None:
        self.dense = DenseRetriever(self.cfg)
        self.bm25 = BM25Retriever(self.cfg)
        self.colbert = None      # optional channel (hybrid_retriever.py:163-169): a failing init only disables it
        if getattr(self.cfg.retrieval, "enable_colbert", False):
            try:
                self.colbert = ColBERTRetriever.from_config(self.cfg)
            except Exception as exc:
                print("[HybridRetriever] ColBERT init failed:", repr(exc))
                traceback.print_exc()
        self._align_key = None   # (store / index mtimes) the alignment verdict below was computed for
        self._aligned = False
        self._graphs: Dict[Any, Any] = {}     # captured single-query pipelines, by (depth, top_k, fusion knobs, store snapshot)
        self._graphs_lock = threading.Lock()
        self.fast_path_used = False
        self.graph = None        # plug a retrieval.GraphRetriever(cfg, graph=<graph store with walk()>) in here: the graph
                                 # store is host-side and injected, its scoring stage runs on the GPU
        self.reranker = None     # optional callable(question, hits) -> hits (cross-encoder rerank is out of scope)

    # ------------------------------------------------------------------ per-channel APIs
    @staticmethod
    def _label(hits: List[RetrievalHit], channel: str, raw_key: Optional[str], replace: bool) -> List[RetrievalHit]:
        """What every per-channel API of the reference does to its hits (hybrid_retriever.py:172-279): order by score,
        rank, source = "retriever", and a breakdown that names the channel (plus the raw score under `raw_key`).  The
        dense channel replaces whatever breakdown the hit carried; the others keep it and fill in what is missing."""
        for h in _by_score(hits):
            h.source = "retriever"
            if replace:
                h.score_breakdown = {"channel": [channel], raw_key: float(h.score)}
                continue
            sb = h.score_breakdown or {}
            sb["channel"] = _as_channel_list(sb.get("channel")) or [channel]
            if raw_key:
                sb.setdefault(raw_key, float(h.score))
            h.score_breakdown = sb
        return hits

    def search_dense(self, question: str, top_k: int = 10) -> List[RetrievalHit]:
        return self._label(self.dense.search(question, max(1, int(top_k))), "dense", "dense_raw", replace=True)

    def search_bm25(self, question: str, top_k: int = 10) -> List[RetrievalHit]:
        pairs = self.bm25.search(question, max(1, int(top_k)))
        return self._label([RetrievalHit(chunk=c, score=float(s), rank=0, source="retriever") for c, s in pairs],
                           "bm25", "bm25_raw", replace=True)

    def search_colbert(self, question: str, top_k: int = 10) -> List[RetrievalHit]:
        if self.colbert is None:
            return []
        try:
            found = [it if isinstance(it, RetrievalHit) else RetrievalHit(chunk=it[0], score=float(it[1]), rank=0, source="retriever")
                     for it in self.colbert.search(question, max(1, int(top_k)))]
            return self._label(found, "colbert", "colbert_raw", replace=False)
        except Exception:        # the reference swallows channel failures (hybrid_retriever.py:233-235)
            return []

    def search_graph(self, question: str, top_k: int = 10, *, decision: Any = None,
                     seeds: Optional[List[RetrievalHit]] = None) -> List[RetrievalHit]:
        if self.graph is None:
            return []
        top_k = max(1, int(top_k))
        if seeds is None:        # no fused list to start from: the head of every channel seeds the walk
            n = int(getattr(self.cfg.retrieval, "graph_seed_k", max(10, top_k * 3)))
            seeds = [h for channel in (self.search_dense, self.search_bm25, self.search_colbert) for h in channel(question, n)[:n]]
        try:
            return self._label(self.graph.search(question, seeds, decision=decision, top_k=top_k), "graph", None, replace=False)
        except Exception:
            return []

    # ------------------------------------------------------------------ main search
    def search(self, question: str, llm: Any = None, top_k: int = 10, decision: Any = None) -> List[RetrievalHit]:
        """hybrid_retriever.py:282-384: oversampled channels -> fusion -> min_final_score filter -> optional graph
        expansion (seeds + graph hits) -> optional rerank of the head -> dedup -> top_k.  `llm` is accepted for the
        reference's positional order; the cross-encoder / LLM reranker itself is injected as `self.reranker`."""
        rcfg = self.cfg.retrieval
        top_k = max(1, int(top_k))
        depth = max(top_k, int(getattr(rcfg, "top_k", top_k * 8) or (top_k * 8)))     # per-channel oversampling
        depth = min(depth, engine.LRAG_MAX_K)                                         # the kernels' list length limit
        top_k = min(top_k, 3 * depth)
        laps = _Laps()
        fast = self._search_graphed(question, top_k, depth, decision, laps)
        if fast is not None:
            return fast
        per_channel = {}
        for name, channel in (("dense", self.search_dense), ("bm25", self.search_bm25), ("colbert", self.search_colbert)):
            per_channel[name] = channel(question, depth)
            laps.mark(name)
        floor = float(getattr(rcfg, "min_final_score", 0.0))
        ranked = [h for h in self._fuse(dense_hits=per_channel["dense"], bm25_hits=per_channel["bm25"],
                                        colbert_hits=per_channel["colbert"]) if float(h.score) >= floor]
        laps.mark("fuse")
        last = "fuse"

        mode = getattr(decision, "mode", None)
        if getattr(rcfg, "enable_graph", False) and mode and str(mode).upper().endswith("GRAPH_AUGMENTED"):
            head = ranked[:int(getattr(rcfg, "graph_seed_k", max(10, top_k * 3)))]
            ranked = head + self.search_graph(question, depth, decision=decision, seeds=head)
            laps.mark("graph")
            last = "graph"

        if getattr(rcfg, "enable_rerank", False) and self.reranker is not None:
            n = int(getattr(rcfg, "rerank_top_n", min(40, max(10, top_k * 4))))
            rescored = self.reranker(question, ranked[:n])
            ranked[:len(rescored)] = rescored
            _by_score(ranked)
            laps.t["rerank_from"] = laps.t[last]
            laps.mark("rerank")

        ranked = _dedup_keep_best(ranked)
        laps.mark("end")
        logger.info("[retrieval] dense=%dms bm25=%dms colbert=%dms fuse=%dms graph=%dms rerank=%dms total=%dms "
                    "enabled(graph=%s,colbert=%s, has_gpu=%s)", laps.ms("start", "dense"), laps.ms("dense", "bm25"),
                    laps.ms("bm25", "colbert"), laps.ms("colbert", "fuse"), laps.ms("fuse", "graph"), laps.ms("rerank_from", "rerank"),
                    laps.ms("start", "end"), int(bool(getattr(rcfg, "enable_graph", False))), int(self.colbert is not None),
                    int(torch.cuda.is_available()))
        return ranked[:top_k]

    # ------------------------------------------------------------------ fusion
    def _fusion_knobs(self):
        rcfg = self.cfg.retrieval
        return dict(method=str(getattr(rcfg, "fusion_method", "rrf_norm_blend")).lower(), rrf_k=int(getattr(rcfg, "rrf_k", 60)),
                    alpha=float(getattr(rcfg, "rrf_alpha", 0.50)), w_dense=float(getattr(rcfg, "dense_weight", 0.55)),
                    w_bm25=float(getattr(rcfg, "bm25_weight", 0.35)), w_colbert=float(getattr(rcfg, "colbert_weight", 0.25)))

    def _fuse(self, *, dense_hits: List[RetrievalHit], bm25_hits: List[RetrievalHit],
              colbert_hits: List[RetrievalHit]) -> List[RetrievalHit]:
        """Same inputs / outputs as the reference's _fuse; the arithmetic runs in liblrag's fusion kernel.
        Equal fused scores are ordered by first appearance (dense, then bm25, then colbert lists); the
        reference orders them by Python set iteration (hybrid_retriever.py:460,484,526)."""
        knobs = self._fusion_knobs()
        lists = {"dense": dense_hits, "bm25": bm25_hits, "colbert": colbert_hits}
        for hs in lists.values():
            hs.sort(key=lambda h: float(h.score), reverse=True)
        if not any(lists.values()):
            return []
        ids: Dict[str, int] = {}
        chunk_of: List[Any] = []
        membership: List[Set[str]] = []
        for ch in CHANNELS:
            for h in lists[ch]:
                j = ids.get(h.chunk.id)
                if j is None:
                    j = ids[h.chunk.id] = len(chunk_of)
                    chunk_of.append(h.chunk)
                    membership.append(set())
                membership[j].add(ch)
        kc = max(1, max(len(v) for v in lists.values()))
        dev = torch.device(getattr(self.cfg, "device", None) or "cuda")
        chans = {}
        for ch in CHANNELS:
            s = np.full((1, kc), engine.PAD_SCORE, dtype=np.float32)
            i = np.full((1, kc), -1, dtype=np.int64)
            seen: Set[int] = set()
            p = 0
            for h in lists[ch]:                    # a chunk listed twice by one channel keeps its best position
                j = ids[h.chunk.id]
                if j in seen:
                    continue
                seen.add(j)
                s[0, p], i[0, p] = float(h.score), j
                p += 1
            chans[ch] = (torch.from_numpy(s).to(dev), torch.from_numpy(i).to(dev))
        n = len(chunk_of)
        fs, fi, bd = engine.fuse_topk(chans["dense"], chans["bm25"], chans["colbert"], k=n, breakdown=True, **knobs)
        fs, fi, bd = fs[0].tolist(), fi[0].tolist(), bd[0].tolist()
        weights = {"dense": knobs["w_dense"], "bm25": knobs["w_bm25"], "colbert": knobs["w_colbert"]}
        out: List[RetrievalHit] = []
        for r, (score, j, b) in enumerate(zip(fs, fi, bd), start=1):
            if j < 0:
                break
            contrib = {"dense": b[5], "bm25": b[6], "colbert": b[7]}
            ch_list = sorted(membership[j], key=lambda c: (float(contrib.get(c, 0.0)), str(c)), reverse=True)
            sb = {"fusion_method": knobs["method"], "rrf_k": knobs["rrf_k"], "alpha": knobs["alpha"], "channel_weights": dict(weights),
                  "channel": ch_list, "channel_contrib": contrib, "rrf_norm": b[0], "weighted_sum": b[1], "dense_norm": b[2],
                  "bm25_norm": b[3], "colbert_norm": b[4]}
            out.append(RetrievalHit(chunk=chunk_of[j], score=float(score), rank=r, source="retriever", score_breakdown=sb))
        return out

    # ------------------------------------------------------------------ shared-row check, single-query fast path
    def _stores_aligned(self) -> bool:
        """True when the dense store, the BM25 index (and the ColBERT meta) list the same chunks in the same order -- what
        scripts.build_index produces -- so that a row number means the same chunk in every channel.  The O(N) comparison
        runs once per store snapshot (keyed by the files' mtimes), not per call."""
        self.dense.store.load()
        self.bm25.load()
        key = (self.dense.store._index_mtime, self.dense.store._meta_mtime, self.bm25._bm25_mtime,
               None if self.colbert is None else self.colbert._meta_mtime)
        if key == self._align_key:
            return self._aligned
        with self._graphs_lock:                   # one thread checks a new snapshot, the others wait for its verdict
            if key == self._align_key:
                return self._aligned
            chunks = self.dense.store.chunks
            ok = len(chunks) == len(self.bm25.chunks) and all(a.id == b.id for a, b in zip(chunks, self.bm25.chunks))
            if ok and self.colbert is not None:
                p2c = self.colbert._pid2chunk
                ok = len(p2c) == len(chunks) and all(p2c.get(i) is not None and p2c[i].id == c.id for i, c in enumerate(chunks))
            self._aligned, self._align_key = ok, key
            self._graphs.clear()                  # captured pipelines hold the old snapshot's device tensors
        return self._aligned

    def _capture_graph(self, key, store, bm, top_k, depth, max_terms, knobs, floor):
        if len(self._graphs) >= 32:
            self._graphs.pop(next(iter(self._graphs)))
        g = self._graphs[key] = engine.GraphedHybridQuery(
            store.index.matrix, bm.device_index, k=top_k, kc=depth, nq=1, max_terms=max_terms, breakdown=True,
            method=knobs["method"], w_dense=knobs["w_dense"], w_bm25=knobs["w_bm25"], w_colbert=knobs["w_colbert"],
            rrf_k=knobs["rrf_k"], alpha=knobs["alpha"], min_final=floor)
        return g

    def _search_graphed(self, question: str, top_k: int, depth: int, decision: Any, laps: "_Laps") -> Optional[List[RetrievalHit]]:
        """search() for the common online case -- dense + BM25 channels over row-aligned stores, no ColBERT channel, no graph
        expansion asked for, no reranker -- as ONE captured CUDA graph per query (engine.GraphedHybridQuery: both scans on
        forked streams, fusion with min_final_score and the breakdown, ~100 us on configs[0]'s corpus) instead of three
        synchronising channel calls and a host-side fusion set-up.  Returns None when the case does not apply; results are
        those of the general path (same kernels; `tests/test_gpu_retrievers.py::test_search_fast_path_equals_general_path`)."""
        rcfg = self.cfg.retrieval
        mode = getattr(decision, "mode", None)
        if (self.colbert is not None or (getattr(rcfg, "enable_rerank", False) and self.reranker is not None)
                or (getattr(rcfg, "enable_graph", False) and self.graph is not None and mode and str(mode).upper().endswith("GRAPH_AUGMENTED"))):
            return None
        if not self._stores_aligned():
            return None
        store, bm = self.dense.store, self.bm25
        tokens = bm.tokenizer(question)
        max_terms = 64
        if len(tokens) > max_terms or store.index.ntotal != len(store.chunks) or store.index.ntotal == 0:
            return None
        knobs = self._fusion_knobs()
        floor = float(getattr(rcfg, "min_final_score", 0.0))
        key = (depth, top_k, tuple(sorted(knobs.items())), floor, id(store.index), id(bm.device_index))
        with self._graphs_lock:               # one thread captures a pipeline, the others wait for it
            g = self._graphs.get(key)
            if g is None:
                g = self._capture_graph(key, store, bm, top_k, depth, max_terms, knobs, floor)
        q_vec = torch.from_numpy(store._embed([question], is_query=True))
        _, qt, _ = bm.host_index.encode_queries([tokens])
        # the captured graph works on static buffers: searches from several host threads (Starlette's thread pool)
        # take turns on it, and everything this call needs is copied out before the next one may start
        with g.lock:
            fs, fi = g.search(q_vec, [qt.tolist()])
            scores, rows, parts = fs[0].tolist(), fi[0].tolist(), g._h_bd[0].tolist()
            in_dense, in_bm25 = set(g._h_di[0].tolist()), set(g._h_bi[0].tolist())
        laps.mark("dense"); laps.mark("bm25"); laps.mark("colbert"); laps.mark("fuse")
        weights = {"dense": knobs["w_dense"], "bm25": knobs["w_bm25"], "colbert": knobs["w_colbert"]}
        chunks = store.chunks
        out: List[RetrievalHit] = []
        for r, (score, j, b) in enumerate(zip(scores, rows, parts), start=1):
            if j < 0:
                break
            contrib = {"dense": b[5], "bm25": b[6], "colbert": b[7]}
            member = [c for c, inside in (("dense", j in in_dense), ("bm25", j in in_bm25)) if inside]
            member.sort(key=lambda c: (float(contrib.get(c, 0.0)), str(c)), reverse=True)
            sb = {"fusion_method": knobs["method"], "rrf_k": knobs["rrf_k"], "alpha": knobs["alpha"], "channel_weights": dict(weights),
                  "channel": member, "channel_contrib": contrib, "rrf_norm": b[0], "weighted_sum": b[1], "dense_norm": b[2],
                  "bm25_norm": b[3], "colbert_norm": b[4]}
            out.append(RetrievalHit(chunk=chunks[j], score=float(score), rank=r, source="retriever", score_breakdown=sb))
        laps.mark("end")
        self.fast_path_used = True
        logger.info("[retrieval] dense=%dms bm25=%dms colbert=%dms fuse=%dms graph=%dms rerank=%dms total=%dms "
                    "enabled(graph=%s,colbert=%s, has_gpu=%s)", laps.ms("start", "dense"), 0, 0, 0, 0, 0, laps.ms("start", "end"),
                    int(bool(getattr(rcfg, "enable_graph", False))), 0, int(torch.cuda.is_available()))
        return out

    # ------------------------------------------------------------------ batched throughput path
    def search_batch(self, questions: Sequence[str], top_k: int = 10) -> List[List[RetrievalHit]]:
        """Every question through every enabled channel in one launch per channel, fused on the device.
        Requires the three stores to index the same chunk list in the same order (what scripts.build_index
        produces); otherwise, or when a reranker is configured, falls back to per-question `search`.  Graph expansion
        needs a routing decision per question (hybrid_retriever.py:315) and is therefore not part of the batch entry
        point; the hits carry the fusion method only, not the per-channel breakdown of `search`."""
        rcfg = self.cfg.retrieval
        top_k = max(1, int(top_k))
        eff = max(top_k, int(getattr(rcfg, "top_k", top_k * 8) or (top_k * 8)))
        chunks = self.dense.store.chunks if self._stores_aligned() else None
        if chunks is None or (getattr(rcfg, "enable_rerank", False) and self.reranker is not None):
            # stores that index different chunk lists, or a reranker to apply: the per-question path does all of it
            return [self.search(q, None, top_k) for q in questions]
        k = min(eff, len(chunks), engine.LRAG_MAX_K)
        store = self.dense.store
        q = torch.from_numpy(store._embed(list(questions), is_query=True))
        dense = store.index.search_device(q, k)
        bm25 = self.bm25.search_ids([self.bm25.tokenizer(x) for x in questions], k)
        colb = self.colbert.search_device(list(questions), k) if self.colbert is not None else None
        knobs = self._fusion_knobs()
        fs, fi = engine.fuse_topk(dense, bm25, colb, k=min(top_k, 3 * k), min_final=float(getattr(rcfg, "min_final_score", 0.0)), **knobs)
        fs, fi = fs.tolist(), fi.tolist()
        out = []
        for rs, ri in zip(fs, fi):
            out.append([RetrievalHit(chunk=chunks[j], score=float(s), rank=r, source="retriever",
                                     score_breakdown={"fusion_method": knobs["method"]})
                        for r, (s, j) in enumerate(zip(rs, ri), start=1) if j >= 0])
        return out
