"""Oracle: BM25-Okapi exactly as the reference executes it.  TEST INFRASTRUCTURE.

parity unpinned: rank-bm25 0.2.2 (requirements.txt:9, pyproject.toml:37) is not
vendored in /root/reference nor installed here; this file restates its published
algorithm and anchors on the reference's call sites:

  * index build   legalrag/retrieval/builders/bm25_builder.py:18-19,39-44
  * scoring       legalrag/retrieval/bm25_retriever.py:73-74  (get_scores)
  * ranking       legalrag/retrieval/bm25_retriever.py:75     (stable full sort)
"""
from __future__ import annotations

import math
import re
from collections import Counter
from typing import Dict, List, Sequence, Tuple

import numpy as np


def tokenize_en(text: str) -> List[str]:
    """Reference English index tokenizer (builders/bm25_builder.py:18-19)."""
    return re.findall(r"[A-Za-z0-9]+(?:'[A-Za-z0-9]+)?", text.lower())


class BM25Okapi:
    """Transcription of rank_bm25.BM25Okapi 0.2.2 (k1=1.5, b=0.75, epsilon=0.25).

    Attribute names match the pickled object the reference stores in bm25.pkl
    (bm25_builder.py:46-51), so a product-side reader can be tested against it.
    """

    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75,
                 epsilon: float = 0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.avgdl = 0.0
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        self.doc_len: List[int] = []
        self.tokenizer = None
        nd: Dict[str, int] = {}
        num_doc = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_doc += len(document)
            frequencies = dict(Counter(document))
            self.doc_freqs.append(frequencies)
            for word in frequencies:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_doc / self.corpus_size
        self._calc_idf(nd)

    def _calc_idf(self, nd: Dict[str, int]) -> None:
        idf_sum = 0.0
        negative_idfs = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative_idfs.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative_idfs:
            self.idf[word] = eps

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:  # every occurrence counts
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


def topk_reference(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """bm25_retriever.py:75 -- ``sorted(range(N), key=scores[i], reverse=True)[:k]``.

    Python's sort is stable and ``reverse=True`` preserves stability, so ties go
    to the LOWER document index; zero-score documents are kept.
    """
    idx = sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)[: int(k)]
    idx = np.asarray(idx, dtype=np.int64)
    return scores[idx].astype(np.float64), idx


def search(bm25: BM25Okapi, tokens: Sequence[str], k: int) -> Tuple[np.ndarray, np.ndarray]:
    return topk_reference(bm25.get_scores(tokens), k)


# ---------------------------------------------------------------------------
# Vectorised restatement for mid-size checks (validated against the literal
# class above in tests/test_oracle.py).  Works on integer term ids.
# ---------------------------------------------------------------------------
class CsrBM25:
    """Same arithmetic as BM25Okapi over a term-major CSR index, float64."""

    def __init__(self, indptr: np.ndarray, doc_id: np.ndarray, tf: np.ndarray, doc_len: np.ndarray,
                 k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.indptr, self.doc_id, self.tf = indptr, doc_id, tf
        self.doc_len = doc_len.astype(np.float64)
        self.N = int(doc_len.shape[0])
        self.k1, self.b = k1, b
        self.avgdl = float(doc_len.sum()) / self.N
        df = np.diff(indptr).astype(np.float64)
        present = df > 0
        idf = np.zeros_like(df)
        idf[present] = np.log(self.N - df[present] + 0.5) - np.log(df[present] + 0.5)
        self.average_idf = float(idf[present].sum() / max(1, int(present.sum())))
        idf[present & (idf < 0)] = epsilon * self.average_idf
        self.idf = idf

    @classmethod
    def from_token_ids(cls, docs: Sequence[np.ndarray], vocab: int, **kw) -> "CsrBM25":
        doc_len = np.array([len(d) for d in docs], dtype=np.int64)
        rows = np.concatenate([np.full(len(d), i, dtype=np.int64) for i, d in enumerate(docs)]) \
            if len(docs) else np.zeros(0, np.int64)
        terms = np.concatenate([np.asarray(d, dtype=np.int64) for d in docs]) if len(docs) else np.zeros(0, np.int64)
        key = terms * len(docs) + rows
        uniq, cnt = np.unique(key, return_counts=True)
        t = uniq // len(docs)
        d = uniq % len(docs)
        indptr = np.zeros(vocab + 1, dtype=np.int64)
        np.add.at(indptr, t + 1, 1)
        indptr = np.cumsum(indptr)
        return cls(indptr, d.astype(np.int64), cnt.astype(np.int64), doc_len, **kw)

    def get_scores(self, term_ids: Sequence[int]) -> np.ndarray:
        score = np.zeros(self.N)
        for t in term_ids:
            if t < 0 or t >= len(self.idf):
                continue
            s, e = self.indptr[t], self.indptr[t + 1]
            d = self.doc_id[s:e]
            f = self.tf[s:e].astype(np.float64)
            score[d] += self.idf[t] * (f * (self.k1 + 1) /
                                       (f + self.k1 * (1 - self.b + self.b * self.doc_len[d] / self.avgdl)))
        return score

    def search(self, term_ids: Sequence[int], k: int) -> Tuple[np.ndarray, np.ndarray]:
        s = self.get_scores(term_ids)
        k = min(int(k), self.N)
        order = np.lexsort((np.arange(self.N), -s))[:k]  # score desc, id asc == stable reverse sort
        return s[order], order.astype(np.int64)
