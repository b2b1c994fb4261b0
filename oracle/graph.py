"""CPU restatement of the scoring stage of the reference's graph-expansion channel.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
/root/reference/legalrag/retrieval/graph_retriever.py:
  _cosine_sim :20-22, _depth_decay :25-27, _relation_weight :30-46, GraphRetriever.search :85-219
(de-duplication by article id :128-136, hydration and the blank-text / language filters :142-157, score =
cosine x depth decay x relation weight x edge confidence :181-192, stable sort and re-rank :211-219).
PINNED: tests/golden/graph_golden.json was produced by executing that code (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

RELATION_WEIGHTS = {"defined_by": 1.20, "defines_term": 1.10, "cite": 1.15, "cited": 1.15, "ref": 1.15, "amend": 1.10,
                    "next": 0.95, "prev": 0.95, "neighbor": 1.00}


def cosine_sim(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.dot(a, b) / ((np.linalg.norm(a) * np.linalg.norm(b)) + 1e-9))


def depth_decay(depth: int, gamma: float = 0.7) -> float:
    d = max(1, int(depth or 1))
    return float(1.0 / ((1.0 + d) ** gamma))


def relation_weight(relations: Sequence[str]) -> float:
    rels = [str(r).lower() for r in (relations or [])]
    if not rels:
        return 1.0
    return float(max(RELATION_WEIGHTS.get(r, 1.0) for r in rels))


def graph_hits(qvec: np.ndarray, nodes: List[Dict[str, Any]], id2row: Dict[str, int], vectors: np.ndarray,
               texts: Sequence[str], langs: Sequence[Optional[str]], *, lang: Optional[str] = None, top_k: int = 10,
               gamma: float = 0.7) -> List[Dict[str, Any]]:
    """nodes: dicts with article_id, graph_depth, relations, meta.  Returns [{row, score, rank, semantic, ...}]."""
    uniq: Dict[str, Dict[str, Any]] = {}
    for n in nodes:
        aid = str(n.get("article_id") or "").strip()
        if aid and aid not in uniq:
            uniq[aid] = n
    out = []
    for aid, n in uniq.items():
        row = id2row.get(aid)
        if row is None or not (texts[row] or "").strip():
            continue
        if lang and (langs[row] or "zh").strip().lower() != lang:
            continue
        gd = int(n.get("graph_depth", 1) or 1)
        conf = float(((n.get("meta") or {}).get("_edge_conf", 1.0)) or 1.0)
        sem = cosine_sim(qvec, vectors[row])
        dd, rw = depth_decay(gd, gamma), relation_weight(n.get("relations") or [])
        out.append({"row": row, "score": float(sem) * dd * rw * conf, "semantic": sem, "depth_decay": dd,
                    "relation_weight": rw, "edge_conf": conf, "graph_depth": gd})
    out.sort(key=lambda h: h["score"], reverse=True)
    for r, h in enumerate(out, start=1):
        h["rank"] = r
    return out[:max(1, int(top_k))]
