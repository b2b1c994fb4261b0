"""Host-side BM25 index: builds the term-major CSR the GPU kernel streams, with rank_bm25.BM25Okapi
statistics (k1 = 1.5, b = 0.75, epsilon = 0.25; idf floor for negative-idf terms) as the reference
builds them at legalrag/retrieval/builders/bm25_builder.py:39-51, and converts queries to term ids.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class Bm25HostIndex:
    indptr: np.ndarray          # [V + 1] int64
    doc_id: np.ndarray          # [nnz] int32, ascending inside each term
    tf: np.ndarray              # [nnz] int32
    doc_len: np.ndarray         # [N] int32
    idf: np.ndarray             # [V] float64 (epsilon floor applied)
    avgdl: float
    average_idf: float
    k1: float = 1.5
    b: float = 0.75
    epsilon: float = 0.25
    term_to_id: Dict[str, int] = field(default_factory=dict)

    @property
    def n_docs(self) -> int:
        return int(self.doc_len.shape[0])

    @property
    def vocab(self) -> int:
        return int(self.indptr.shape[0] - 1)

    # ------------------------------------------------------------------ builders
    @classmethod
    def from_postings(cls, indptr, doc_id, tf, doc_len, *, k1=1.5, b=0.75, epsilon=0.25, term_to_id=None,
                      n_docs_global: Optional[int] = None, df_global: Optional[np.ndarray] = None,
                      avgdl_global: Optional[float] = None) -> "Bm25HostIndex":
        """idf / avgdl come from the GLOBAL corpus statistics when a shard is being built."""
        indptr = np.asarray(indptr, dtype=np.int64)
        doc_len = np.asarray(doc_len, dtype=np.int32)
        N = int(n_docs_global if n_docs_global is not None else doc_len.shape[0])
        df = np.asarray(df_global if df_global is not None else np.diff(indptr), dtype=np.float64)
        avgdl = float(avgdl_global if avgdl_global is not None else (doc_len.sum(dtype=np.int64) / max(1, N)))
        present = df > 0
        idf = np.zeros_like(df)
        idf[present] = np.log(N - df[present] + 0.5) - np.log(df[present] + 0.5)
        # rank_bm25 sums the idf values left to right in vocabulary insertion order; pairwise numpy
        # summation differs from that only in the last bits
        average_idf = float(idf[present].sum() / max(1, int(present.sum())))
        idf[present & (idf < 0)] = epsilon * average_idf
        return cls(indptr, np.asarray(doc_id, dtype=np.int32), np.asarray(tf, dtype=np.int32), doc_len, idf, avgdl,
                   average_idf, k1, b, epsilon, dict(term_to_id or {}))

    @classmethod
    def from_token_ids(cls, docs: Sequence[np.ndarray], vocab: int, **kw) -> "Bm25HostIndex":
        n = len(docs)
        doc_len = np.array([len(d) for d in docs], dtype=np.int32)
        if n == 0 or int(doc_len.sum()) == 0:
            return cls.from_postings(np.zeros(vocab + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), doc_len, **kw)
        rows = np.repeat(np.arange(n, dtype=np.int64), doc_len)
        terms = np.concatenate([np.asarray(d, dtype=np.int64) for d in docs])
        uniq, cnt = np.unique(terms * n + rows, return_counts=True)
        t, d = uniq // n, uniq % n
        indptr = np.zeros(vocab + 1, dtype=np.int64)
        np.add.at(indptr, t + 1, 1)
        return cls.from_postings(np.cumsum(indptr), d, cnt, doc_len, **kw)

    @classmethod
    def from_tokens(cls, corpus_tokens: Sequence[Sequence[str]], **kw) -> "Bm25HostIndex":
        """corpus_tokens as BM25Okapi(corpus_tokens) takes them (bm25_builder.py:44)."""
        term_to_id: Dict[str, int] = {}
        docs = []
        for doc in corpus_tokens:
            ids = np.empty(len(doc), dtype=np.int64)
            for j, w in enumerate(doc):
                i = term_to_id.get(w)
                if i is None:
                    i = term_to_id[w] = len(term_to_id)
                ids[j] = i
            docs.append(ids)
        return cls.from_token_ids(docs, len(term_to_id), term_to_id=term_to_id, **kw)

    @classmethod
    def from_okapi(cls, bm25) -> "Bm25HostIndex":
        """From a (possibly shim-unpickled) rank_bm25.BM25Okapi: uses ITS idf / avgdl verbatim, so scores
        follow whatever library version wrote the pickle (bm25_retriever.py:47-66)."""
        doc_freqs: List[Dict[str, int]] = list(bm25.doc_freqs)
        term_to_id: Dict[str, int] = {}
        for w in bm25.idf:
            term_to_id[w] = len(term_to_id)
        t_list, d_list, f_list = [], [], []
        for d, fr in enumerate(doc_freqs):
            for w, f in fr.items():
                i = term_to_id.get(w)
                if i is None:
                    i = term_to_id[w] = len(term_to_id)
                t_list.append(i); d_list.append(d); f_list.append(int(f))
        V = len(term_to_id)
        t = np.asarray(t_list, dtype=np.int64); d = np.asarray(d_list, dtype=np.int64); f = np.asarray(f_list, dtype=np.int32)
        order = np.lexsort((d, t))
        indptr = np.zeros(V + 1, dtype=np.int64)
        np.add.at(indptr, t + 1, 1)
        idf = np.zeros(V, dtype=np.float64)
        for w, v in bm25.idf.items():
            idf[term_to_id[w]] = float(v)
        doc_len = np.asarray(bm25.doc_len, dtype=np.int32)
        return cls(np.cumsum(indptr), d[order].astype(np.int32), f[order], doc_len, idf, float(bm25.avgdl),
                   float(getattr(bm25, "average_idf", 0.0)), float(bm25.k1), float(bm25.b),
                   float(getattr(bm25, "epsilon", 0.25)), term_to_id)

    # ------------------------------------------------------------------ scoring tables
    def impacts(self) -> np.ndarray:
        """fp32 per-posting score contribution, computed in fp64 exactly as get_scores does per term."""
        t_of = np.repeat(np.arange(self.vocab, dtype=np.int64), np.diff(self.indptr))
        f = self.tf.astype(np.float64)
        dl = self.doc_len[self.doc_id].astype(np.float64)
        imp = self.idf[t_of] * (f * (self.k1 + 1) / (f + self.k1 * (1 - self.b + self.b * dl / self.avgdl)))
        return imp.astype(np.float32)

    def to_device(self, device, lo: int = 0, hi: Optional[int] = None):
        """Device-resident shard holding documents [lo, hi) with LOCAL doc ids and id_base = lo."""
        import torch
        from .engine import Bm25DeviceIndex
        hi = self.n_docs if hi is None else hi
        imp = self.impacts()
        # one bound for every shard of the corpus: the fixed-point scale of a query, and with it every score bit,
        # is then the same whichever way the corpus is sharded
        bound = float(np.abs(imp).max()) if imp.size else 0.0
        if lo == 0 and hi == self.n_docs:
            indptr, doc_id = self.indptr, self.doc_id
        else:
            keep = (self.doc_id >= lo) & (self.doc_id < hi)
            t_of = np.repeat(np.arange(self.vocab, dtype=np.int64), np.diff(self.indptr))
            indptr = np.zeros(self.vocab + 1, dtype=np.int64)
            np.add.at(indptr, t_of[keep] + 1, 1)
            indptr = np.cumsum(indptr)
            doc_id = (self.doc_id[keep] - lo).astype(np.int32)
            imp = imp[keep]
        nonneg = bool((imp >= 0).all()) if imp.size else True
        return Bm25DeviceIndex(torch.from_numpy(np.ascontiguousarray(indptr)).to(device),
                               torch.from_numpy(np.ascontiguousarray(doc_id)).to(device),
                               torch.from_numpy(np.ascontiguousarray(imp)).to(device), hi - lo, nonneg, lo, bound)

    # ------------------------------------------------------------------ queries
    def encode_queries(self, queries: Sequence[Sequence]) -> Tuple[np.ndarray, np.ndarray, int]:
        """Token lists (str, or int term ids) -> (q_indptr int64 [nq+1], q_term int32, max tokens per query).
        Unknown tokens become -1; repeats are kept (each occurrence scores, as in get_scores)."""
        lens = np.array([len(q) for q in queries], dtype=np.int64)
        q_indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        flat = np.full(int(lens.sum()), -1, dtype=np.int32)
        p = 0
        for q in queries:
            for w in q:
                if isinstance(w, str):
                    flat[p] = self.term_to_id.get(w, -1)
                else:
                    wi = int(w)
                    flat[p] = wi if 0 <= wi < self.vocab else -1
                p += 1
        return q_indptr, flat, int(lens.max()) if len(lens) else 0
