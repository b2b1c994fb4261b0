"""GraphRetriever: the scoring stage of the reference's graph-expansion channel on the GPU
(replaces legalrag/retrieval/graph_retriever.py:51-219).

The graph store itself (JSONL loading, BFS walk; legalrag/retrieval/graph_store.py) stays host-side and is
injected: any object with `walk(start_ids=, relation_max_depth=, limit=, rel_types=, min_conf=)` returning
nodes with `article_id`, `graph_depth`, `relations`, `meta` -- e.g. the reference's own LawGraphStore.

What changes under the hood: the reference embeds the <= graph_limit neighbour texts AGAIN with the dense
encoder and takes cosines one by one (:177-186).  Neighbours are hydrated from the vector store's chunks, so
their unit-norm embeddings are already rows of the dense corpus in HBM: one gathered inner-product kernel
(`lrag_dense_gather_scores_bf16`) scores them all, and only the question is encoded.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .. import engine
from ..schemas import LawChunk, RetrievalHit
from .vector_store import VectorStore

RELATION_WEIGHTS = {"defined_by": 1.20, "defines_term": 1.10, "cite": 1.15, "cited": 1.15, "ref": 1.15, "amend": 1.10,
                    "next": 0.95, "prev": 0.95, "neighbor": 1.00}           # graph_retriever.py:34-44


def _depth_decay(depth: int, gamma: float = 0.7) -> float:                  # :25-27
    d = max(1, int(depth or 1))
    return float(1.0 / ((1.0 + d) ** gamma))


def _relation_weight(relations: List[str]) -> float:                        # :30-46
    rels = [str(r).lower() for r in (relations or [])]
    if not rels:
        return 1.0
    return float(max(RELATION_WEIGHTS.get(r, 1.0) for r in rels))


@dataclass
class GraphRetriever:
    cfg: Any
    graph: Any = None
    store: Optional[VectorStore] = None
    id2chunk: Optional[Dict[str, LawChunk]] = None

    def __post_init__(self) -> None:
        if self.store is None:
            self.store = VectorStore.from_config(self.cfg)
        self.store.load()
        self._snapshot = None
        self._id2row: Dict[str, int] = {}
        self._refresh()

    def _refresh(self) -> None:
        """article id -> (chunk, corpus row); rebuilt when the store swapped its snapshot in (reload / incremental add)."""
        chunks = getattr(self.store, "chunks", None) or []
        if self._snapshot == (id(chunks), len(chunks)):
            return
        self.id2chunk, self._id2row = {}, {}
        for row, c in enumerate(chunks):
            aid = getattr(c, "article_id", None) or getattr(c, "id", None)
            if aid:
                self.id2chunk[str(aid)] = c        # later duplicates win, like the reference's dict fill (:76-80)
                self._id2row[str(aid)] = row
        self._snapshot = (id(chunks), len(chunks))

    def search(self, question: str, seeds: List[Any], *, decision: Any = None, lang: Optional[str] = None,
               top_k: int = 10) -> List[RetrievalHit]:
        if self.graph is None:
            return []
        self._refresh()
        rcfg = getattr(self.cfg, "retrieval", None)
        eff_top_k = max(1, int(top_k))
        relation_depths = rcfg.graph_walk_depths if hasattr(rcfg, "graph_walk_depths") else {"default": 2}
        limit = int(getattr(rcfg, "graph_limit", eff_top_k * 8) if rcfg else eff_top_k * 8)
        rel_types = getattr(rcfg, "graph_rel_types", None) if rcfg else None
        min_conf = float(getattr(rcfg, "graph_min_conf", 0.0) if rcfg else 0.0)
        gamma = float(getattr(rcfg, "graph_depth_gamma", 0.7) if rcfg else 0.7)

        seed_ids: List[str] = []
        for h in seeds or []:
            c = getattr(h, "chunk", None)
            if c is None:
                continue
            aid = getattr(c, "article_id", None) or getattr(c, "id", None)
            if aid:
                seed_ids.append(str(aid))
        if not seed_ids:
            return []
        nodes = self.graph.walk(start_ids=seed_ids, relation_max_depth=relation_depths, limit=limit, rel_types=rel_types,
                                min_conf=min_conf)
        if not nodes:
            return []
        uniq: Dict[str, Any] = {}
        for n in nodes:                                                     # :128-136
            aid = str(getattr(n, "article_id", "") or "").strip()
            if aid and aid not in uniq:
                uniq[aid] = n

        rows, kept, meta = [], [], []
        for aid, n in uniq.items():                                         # :142-172
            c = self.id2chunk.get(aid)
            if not c or not (getattr(c, "text", "") or "").strip():
                continue
            if lang and (getattr(c, "lang", None) or "zh").strip().lower() != lang:
                continue
            cc = copy.copy(c)
            try:
                setattr(cc, "source", "graph")
            except Exception:
                pass
            kept.append(cc)
            rows.append(self._id2row[aid])
            meta.append((int(getattr(n, "graph_depth", 1) or 1), list(getattr(n, "relations", []) or []),
                         float(((getattr(n, "meta", {}) or {}).get("_edge_conf", 1.0)) or 1.0)))
        if not kept:
            return []

        # the reference embeds the question as a passage here (no query instruction, :177) and divides by the
        # norms (:20-22); corpus rows are unit-norm already, the question vector is normalised on the host
        qvec = np.asarray(self.store._embed([question]), dtype=np.float32).reshape(-1)
        qvec = qvec / (float(np.linalg.norm(qvec)) + 1e-9)
        X = self.store.index.matrix
        Q = torch.from_numpy(qvec[None, :]).to(X.device).to(torch.bfloat16)
        sem = engine.gather_scores(X, Q, torch.tensor([rows], dtype=torch.int64, device=X.device))[0].cpu().numpy()

        hits: List[RetrievalHit] = []
        for i, (c, s, (gd, rels, conf)) in enumerate(zip(kept, sem, meta), start=1):
            dd, rw = _depth_decay(gd, gamma=gamma), _relation_weight(rels)
            final = float(s) * dd * rw * conf
            hits.append(RetrievalHit(chunk=c, score=final, rank=i, source="graph",
                                     score_breakdown={"channel": "graph", "semantic": float(s), "depth_decay": dd,
                                                      "relation_weight": rw, "edge_conf": conf, "final": final,
                                                      "graph_depth": gd, "relations": rels}))
        hits.sort(key=lambda h: float(h.score or 0.0), reverse=True)        # stable, like :211
        for r, h in enumerate(hits, start=1):
            h.rank = r
        return hits[:eff_top_k]
