"""DenseRetriever (replaces legalrag/retrieval/dense_retriever.py:13-60)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from ..schemas import RetrievalHit
from .vector_store import VectorStore


@dataclass
class DenseRetriever:
    cfg: object
    store: Optional[VectorStore] = None

    def __post_init__(self) -> None:
        if self.store is None:
            self.store = VectorStore.from_config(self.cfg)

    def search(self, query: str, top_k: int) -> List[RetrievalHit]:
        self.store.load()
        k = max(1, int(top_k))
        q_vec = self.store._embed([query], is_query=True)
        scores, idxs = self.store.index.search(q_vec, k)
        return self._hits(scores[0].tolist(), idxs[0].tolist())

    def _hits(self, scores, idxs) -> List[RetrievalHit]:
        hits: List[RetrievalHit] = []
        for rank, (i, s) in enumerate(zip(idxs, scores), start=1):
            if i < 0 or i >= len(self.store.chunks):
                continue
            hits.append(RetrievalHit(chunk=self.store.chunks[i], score=float(s), rank=rank, source="retriever",
                                     semantic_score=float(s)))
        return hits

    def search_batch(self, queries: Sequence[str], top_k: int) -> List[List[RetrievalHit]]:
        """All queries in one scan (the reference embeds and searches one query per call)."""
        self.store.load()
        k = max(1, int(top_k))
        q = self.store._embed(list(queries), is_query=True)
        scores, idxs = self.store.index.search(q, k)
        return [self._hits(s.tolist(), i.tolist()) for s, i in zip(scores, idxs)]
