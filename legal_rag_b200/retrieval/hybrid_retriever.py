"""HybridRetriever (replaces legalrag/retrieval/hybrid_retriever.py:133-551): same channels, same four
fusion methods, same score_breakdown, with every scoring step on the GPU.  `search(question, llm, top_k,
decision)` keeps the reference's positional order; `search_batch` is the throughput entry point (all
queries through every channel in one kernel launch each, fused on the device)."""
from __future__ import annotations

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
    if x is None:
        return []
    if isinstance(x, (list, set, tuple)):
        return [str(i) for i in x]
    return [str(x)]


def _dedup_keep_best(hits: List[RetrievalHit]) -> List[RetrievalHit]:
    """Best-scoring hit per chunk.id, channels united, channel_contrib summed, re-ranked
    (hybrid_retriever.py:71-130)."""
    best: Dict[str, RetrievalHit] = {}
    for h in hits:
        cid = h.chunk.id
        sb = h.score_breakdown or {}
        if cid not in best:
            if "channel" in sb:
                sb["channel"] = _as_channel_list(sb.get("channel"))
                h.score_breakdown = sb
            best[cid] = h
            continue
        b = best[cid]
        sb_best = b.score_breakdown or {}
        merged_channels = list(set(_as_channel_list(sb_best.get("channel"))) | set(_as_channel_list(sb.get("channel"))))
        merged_contrib: Dict[str, float] = {}
        for src in (sb_best.get("channel_contrib", {}) or {}, sb.get("channel_contrib", {}) or {}):
            if isinstance(src, dict):
                for k, v in src.items():
                    merged_contrib[str(k)] = merged_contrib.get(str(k), 0.0) + float(v)
        if float(h.score) > float(b.score):
            best[cid] = h
        rep = best[cid]
        sb_rep = rep.score_breakdown or {}
        if merged_contrib:
            merged_channels.sort(key=lambda c: float(merged_contrib.get(c, 0.0)), reverse=True)
            sb_rep["channel_contrib"] = merged_contrib
        else:
            merged_channels.sort()
        sb_rep["channel"] = merged_channels
        rep.score_breakdown = sb_rep
    out = list(best.values())
    out.sort(key=lambda x: float(x.score), reverse=True)
    for i, h in enumerate(out, start=1):
        h.rank = i
    return out


@dataclass
class HybridRetriever:
    cfg: Any

    def __post_init__(self) -> None:
        self.dense = DenseRetriever(self.cfg)
        self.bm25 = BM25Retriever(self.cfg)
        self.colbert = None
        if getattr(self.cfg.retrieval, "enable_colbert", False):
            try:
                self.colbert = ColBERTRetriever.from_config(self.cfg)
            except Exception as e:                       # hybrid_retriever.py:163-169: the channel is optional
                print("[HybridRetriever] ColBERT init failed:", repr(e))
                traceback.print_exc()
                self.colbert = None
        self.graph = None        # plug a retrieval.GraphRetriever(cfg, graph=<graph store with walk()>) in here: the graph
                                 # store is host-side and injected, its scoring stage runs on the GPU
        self.reranker = None     # optional callable(question, hits) -> hits (cross-encoder rerank is out of scope)

    # ------------------------------------------------------------------ per-channel APIs
    def search_dense(self, question: str, top_k: int = 10) -> List[RetrievalHit]:
        top_k = max(1, int(top_k))
        hits = self.dense.search(question, top_k)
        hits.sort(key=lambda h: float(h.score), reverse=True)
        for i, h in enumerate(hits, start=1):
            h.rank = i
            h.source = "retriever"
            h.score_breakdown = {"channel": ["dense"], "dense_raw": float(h.score)}
        return hits

    def search_bm25(self, question: str, top_k: int = 10) -> List[RetrievalHit]:
        top_k = max(1, int(top_k))
        hits = [RetrievalHit(chunk=c, score=float(s), rank=i, source="retriever",
                             score_breakdown={"channel": ["bm25"], "bm25_raw": float(s)})
                for i, (c, s) in enumerate(self.bm25.search(question, top_k), start=1)]
        hits.sort(key=lambda h: float(h.score), reverse=True)
        for i, h in enumerate(hits, start=1):
            h.rank = i
        return hits

    def search_colbert(self, question: str, top_k: int = 10) -> List[RetrievalHit]:
        top_k = max(1, int(top_k))
        if self.colbert is None:
            return []
        try:
            hits: List[RetrievalHit] = []
            for item in self.colbert.search(question, top_k):
                if isinstance(item, RetrievalHit):
                    hits.append(item)
                else:
                    c, s = item
                    hits.append(RetrievalHit(chunk=c, score=float(s), rank=0, source="retriever",
                                             score_breakdown={"channel": ["colbert"], "colbert_raw": float(s)}))
            hits.sort(key=lambda h: float(h.score), reverse=True)
            for i, h in enumerate(hits, start=1):
                h.rank = i
                h.source = "retriever"
                sb = h.score_breakdown or {}
                sb["channel"] = _as_channel_list(sb.get("channel")) or ["colbert"]
                sb.setdefault("colbert_raw", float(h.score))
                h.score_breakdown = sb
            return hits
        except Exception:
            return []

    def search_graph(self, question: str, top_k: int = 10, *, decision: Any = None,
                     seeds: Optional[List[RetrievalHit]] = None) -> List[RetrievalHit]:
        top_k = max(1, int(top_k))
        if self.graph is None:
            return []
        if seeds is None:
            seed_n = int(getattr(self.cfg.retrieval, "graph_seed_k", max(10, top_k * 3)))
            seeds = self.search_dense(question, seed_n)[:seed_n] + self.search_bm25(question, seed_n)[:seed_n] \
                + self.search_colbert(question, seed_n)[:seed_n]
        try:
            hits = self.graph.search(question, seeds, decision=decision, top_k=top_k)
            hits.sort(key=lambda h: float(h.score), reverse=True)
            for i, h in enumerate(hits, start=1):
                h.rank = i
                h.source = "retriever"
                sb = h.score_breakdown or {}
                sb["channel"] = _as_channel_list(sb.get("channel")) or ["graph"]
                h.score_breakdown = sb
            return hits
        except Exception:
            return []

    # ------------------------------------------------------------------ main search
    def search(self, question: str, llm: Any = None, top_k: int = 10, decision: Any = None) -> List[RetrievalHit]:
        rcfg = self.cfg.retrieval
        top_k = max(1, int(top_k))
        t_start = time.time()
        eff_top_k = int(getattr(rcfg, "top_k", top_k * 8) or (top_k * 8))
        if eff_top_k < top_k:
            eff_top_k = top_k

        t0 = time.time()
        dense_hits = self.search_dense(question, eff_top_k)
        t1 = time.time()
        bm25_hits = self.search_bm25(question, eff_top_k)
        t2 = time.time()
        colbert_hits = self.search_colbert(question, eff_top_k)
        t3 = time.time()
        fused = self._fuse(dense_hits=dense_hits, bm25_hits=bm25_hits, colbert_hits=colbert_hits)
        t4 = time.time()

        min_final = float(getattr(rcfg, "min_final_score", 0.0))
        fused = [h for h in fused if float(h.score) >= min_final]

        t_graph = None
        mode = getattr(decision, "mode", None)
        if getattr(rcfg, "enable_graph", False) and mode and (str(mode).upper().endswith("GRAPH_AUGMENTED")):
            seed_n = int(getattr(rcfg, "graph_seed_k", max(10, top_k * 3)))
            seeds = fused[:seed_n]
            fused = seeds + self.search_graph(question, eff_top_k, decision=decision, seeds=seeds)
            t_graph = time.time()

        t_rerank = None
        if getattr(rcfg, "enable_rerank", False) and self.reranker is not None:
            rerank_top_n = int(getattr(rcfg, "rerank_top_n", min(40, max(10, top_k * 4))))
            new_hits = self.reranker(question, fused[:rerank_top_n])
            fused[:len(new_hits)] = new_hits
            fused.sort(key=lambda x: float(x.score), reverse=True)
            for i, h in enumerate(fused, start=1):
                h.rank = i
            t_rerank = time.time()

        fused = _dedup_keep_best(fused)
        t_end = time.time()
        ms = lambda a, b: int((b - a) * 1000)   # noqa: E731
        logger.info("[retrieval] dense=%dms bm25=%dms colbert=%dms fuse=%dms graph=%dms rerank=%dms total=%dms "
                    "enabled(graph=%s,colbert=%s, has_gpu=%s)", ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4),
                    ms(t4, t_graph) if t_graph else 0, ms((t_graph or t4), t_rerank) if t_rerank else 0, ms(t_start, t_end),
                    int(bool(getattr(rcfg, "enable_graph", False))), int(self.colbert is not None), int(torch.cuda.is_available()))
        return fused[:top_k]

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

    # ------------------------------------------------------------------ batched throughput path
    def search_batch(self, questions: Sequence[str], top_k: int = 10) -> List[List[RetrievalHit]]:
        """Every question through every enabled channel in one launch per channel, fused on the device.
        Requires the three stores to index the same chunk list in the same order (what scripts.build_index
        produces); otherwise falls back to per-question `search`."""
        rcfg = self.cfg.retrieval
        top_k = max(1, int(top_k))
        eff = max(top_k, int(getattr(rcfg, "top_k", top_k * 8) or (top_k * 8)))
        self.dense.store.load()
        self.bm25.load()
        chunks = self.dense.store.chunks
        aligned = len(chunks) == len(self.bm25.chunks) and all(a.id == b.id for a, b in zip(chunks, self.bm25.chunks))
        if aligned and self.colbert is not None:
            p2c = self.colbert._pid2chunk
            aligned = len(p2c) == len(chunks) and all(p2c.get(i) is not None and p2c[i].id == c.id for i, c in enumerate(chunks))
        if not aligned:
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
