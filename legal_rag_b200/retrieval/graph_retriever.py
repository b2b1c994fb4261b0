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

    @staticmethod
    def _article_id(obj: Any) -> str:
        return str(getattr(obj, "article_id", None) or getattr(obj, "id", None) or "")

    def _walk_params(self, top_k: int) -> Dict[str, Any]:
        """The knobs of the graph walk and of the depth decay, with the reference's defaults (graph_retriever.py:98-108)."""
        rcfg = getattr(self.cfg, "retrieval", None)
        pick = lambda name, default: getattr(rcfg, name, default) if rcfg is not None else default      # noqa: E731
        return {"depths": rcfg.graph_walk_depths if hasattr(rcfg, "graph_walk_depths") else {"default": 2},
                "limit": int(pick("graph_limit", top_k * 8)), "rel_types": pick("graph_rel_types", None),
                "min_conf": float(pick("graph_min_conf", 0.0)), "gamma": float(pick("graph_depth_gamma", 0.7))}

    def search(self, question: str, seeds: List[Any], *, decision: Any = None, lang: Optional[str] = None,
               top_k: int = 10) -> List[RetrievalHit]:
        """Walk the graph from the seeds' articles, keep the neighbours the store knows (first visit of an article wins,
        empty texts and other languages are dropped), score them against the question in one gathered inner-product
        launch, weight by depth decay x relation weight x edge confidence, best first (graph_retriever.py:85-219)."""
        if self.graph is None:
            return []
        self._refresh()
        depth_k = max(1, int(top_k))
        knobs = self._walk_params(depth_k)
        start = [aid for aid in (self._article_id(c) for c in (getattr(h, "chunk", None) for h in (seeds or [])) if c is not None)
                 if aid]
        if not start:
            return []
        visited = self.graph.walk(start_ids=start, relation_max_depth=knobs["depths"], limit=knobs["limit"],
                                  rel_types=knobs["rel_types"], min_conf=knobs["min_conf"])
        first_visit: Dict[str, Any] = {}
        for node in visited or []:
            aid = str(getattr(node, "article_id", "") or "").strip()
            if aid:
                first_visit.setdefault(aid, node)

        rows: List[int] = []
        neighbours: List[tuple] = []           # (chunk copy tagged source="graph", depth, relations, edge confidence)
        for aid, node in first_visit.items():
            known = self.id2chunk.get(aid)
            if not known or not (getattr(known, "text", "") or "").strip():
                continue
            if lang and (getattr(known, "lang", None) or "zh").strip().lower() != lang:
                continue
            tagged = copy.copy(known)
            try:
                tagged.source = "graph"
            except Exception:       # frozen models keep their source
                pass
            edge_conf = float(((getattr(node, "meta", {}) or {}).get("_edge_conf", 1.0)) or 1.0)
            neighbours.append((tagged, int(getattr(node, "graph_depth", 1) or 1), list(getattr(node, "relations", []) or []), edge_conf))
            rows.append(self._id2row[aid])
        if not neighbours:
            return []

        # the reference embeds the question as a passage here (no query instruction, :177) and divides by the
        # norms (:20-22); corpus rows are unit-norm already, the question vector is normalised on the host
        qvec = np.asarray(self.store._embed([question]), dtype=np.float32).reshape(-1)
        qvec = qvec / (float(np.linalg.norm(qvec)) + 1e-9)
        X = self.store.index.matrix
        Q = torch.from_numpy(qvec[None, :]).to(X.device).to(torch.bfloat16)
        cosines = engine.gather_scores(X, Q, torch.tensor([rows], dtype=torch.int64, device=X.device))[0].cpu().tolist()

        hits: List[RetrievalHit] = []
        for cos, (chunk, depth, relations, edge_conf) in zip(cosines, neighbours):
            decay, rel_w = _depth_decay(depth, gamma=knobs["gamma"]), _relation_weight(relations)
            value = float(cos) * decay * rel_w * edge_conf
            hits.append(RetrievalHit(chunk=chunk, score=value, rank=0, source="graph",
                                     score_breakdown={"channel": "graph", "semantic": float(cos), "depth_decay": decay,
                                                      "relation_weight": rel_w, "edge_conf": edge_conf, "final": value,
                                                      "graph_depth": depth, "relations": relations}))
        hits.sort(key=lambda h: -float(h.score or 0.0))                     # stable, like the reference's sort (:211)
        for pos, h in enumerate(hits):
            h.rank = pos + 1
        return hits[:depth_k]
