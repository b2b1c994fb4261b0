"""Writers of the index artifacts (the reference's legalrag/retrieval/builders/*), producing files the
reference can read back: faiss.index + faiss_meta.jsonl, bm25.pkl, colbert_meta.jsonl (+ this engine's
token store).  The encoders are pluggable (retrieval/encoders.py)."""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

from ..schemas import LawChunk
from . import artifacts, encoders
from .colbert_retriever import build_token_store as build_colbert_index  # noqa: F401
from .vector_store import VectorStore


def build_faiss_index(cfg, chunks: Sequence[LawChunk], encoder=None, flat: bool = True) -> None:
    """builders/faiss_builder.py:66-104.  The reference writes IndexHNSWFlat; an exact engine needs only the
    flat storage, so by default a plain IndexFlatIP file is written (the reference's VectorStore.load opens
    it unchanged: vector_store.py:112-117).  flat=False writes the reference's own type, IndexHNSWFlat, with a single-level
    graph of exact nearest-neighbour links (computed by the dense scan kernel when a GPU is present)."""
    rcfg = cfg.retrieval
    enc = encoder or encoders.make_dense_encoder(str(rcfg.embedding_model), "cpu")
    X = np.asarray(enc.encode([c.text for c in chunks], batch_size=64, max_length=512), dtype=np.float32)
    Path(rcfg.faiss_index_file).parent.mkdir(parents=True, exist_ok=True)
    Path(rcfg.faiss_meta_file).parent.mkdir(parents=True, exist_ok=True)
    if flat:
        artifacts.write_faiss_flat(rcfg.faiss_index_file, X)
    else:
        M = int(getattr(rcfg, "hnsw_m", 64))
        artifacts.write_faiss_hnsw_flat(rcfg.faiss_index_file, X, M, int(getattr(rcfg, "hnsw_ef_construction", 400)),
                                        int(getattr(rcfg, "hnsw_ef_search", 512)), neighbors=_knn_links(X, 2 * M))
    artifacts.write_meta_jsonl(rcfg.faiss_meta_file, chunks)


def _knn_links(X: np.ndarray, m: int) -> Optional[np.ndarray]:
    """Exact inner-product nearest neighbours of every row (self excluded) through the dense scan kernel; None without a GPU
    (the writer then falls back to numpy for small corpora)."""
    import torch
    n = X.shape[0]
    if not torch.cuda.is_available() or n < 2:
        return None
    from .. import engine
    k = min(m + 1, n, engine.LRAG_MAX_K)
    Xd = torch.from_numpy(X).cuda().to(torch.bfloat16)
    out = np.full((n, m), -1, dtype=np.int32)
    for lo in range(0, n, 4096):
        _, idx = engine.dense_topk(Xd, Xd[lo:lo + 4096], k)
        idx = idx.cpu().numpy()
        for r in range(idx.shape[0]):
            row = idx[r][(idx[r] != lo + r) & (idx[r] >= 0)][:m]
            out[lo + r, :len(row)] = row
    return out


def build_bm25_index(cfg, chunks: Sequence[LawChunk], tokenizer=None) -> None:
    """builders/bm25_builder.py:22-53: English chunks use the lower-casing regex, others jieba."""
    rcfg = cfg.retrieval
    lang = (getattr(chunks[0], "lang", None) or "zh").strip().lower() if chunks else "zh"
    if tokenizer is None:
        tokenizer = encoders.tokenize_en if lang == "en" else encoders.default_query_tokenizer()
    state = artifacts.okapi_state_from_tokens([tokenizer(c.text) for c in chunks])
    Path(rcfg.bm25_index_file).parent.mkdir(parents=True, exist_ok=True)
    artifacts.write_bm25_pickle(rcfg.bm25_index_file, state, chunks)


def _load_jsonl_chunks(path) -> List[LawChunk]:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(path)
    chunks: List[LawChunk] = []
    with path.open("r", encoding="utf-8") as f:
        for line in f:
            if line.strip():
                chunks.append(LawChunk.model_validate(json.loads(line)))
    return chunks


def _file_lock(path):
    """The reference serialises writers with filelock.FileLock (incremental_dense_builder.py:45-46)."""
    try:
        from filelock import FileLock
        return FileLock(str(path))
    except ImportError:      # single-writer deployments work without it
        import contextlib
        return contextlib.nullcontext()


class IncrementalDenseBuilder:
    """builders/incremental_dense_builder.py:17-78: append the chunks whose id is not indexed yet.  The meta
    file is appended BEFORE the index file is replaced (atomically), so a reader never sees an index row
    without its chunk; writers are serialised by a lock file next to the index."""

    def __init__(self, cfg, store: Optional[VectorStore] = None):
        self.cfg = cfg
        self.store = store or VectorStore.from_config(cfg)
        self.vs = self.store          # the reference's attribute name

    def add_chunks(self, chunks: Sequence[LawChunk]) -> int:
        if not chunks:
            return 0
        lock_path = Path(self.cfg.retrieval.faiss_index_file).with_suffix(".lock")
        with _file_lock(lock_path):
            self.store.load()
            exist_ids = {c.id for c in self.store.chunks}
            new_chunks, seen = [], set()
            for c in chunks:
                if c.id not in exist_ids and c.id not in seen:
                    new_chunks.append(c)
                    seen.add(c.id)
            if not new_chunks:
                return 0
            vecs = self.store._embed([c.text for c in new_chunks], is_query=False)
            self.store.meta_path.parent.mkdir(parents=True, exist_ok=True)
            with open(self.store.meta_path, "a", encoding="utf-8") as f:
                for c in new_chunks:
                    f.write(json.dumps(c.model_dump(), ensure_ascii=False) + "\n")
            # The file keeps the ORIGINAL fp32 rows (the reference appends fp32 vectors and rewrites its own index object,
            # incremental_dense_builder.py:60-75): they are read back from the file, not from the bf16 device matrix, so
            # repeated increments never degrade what the reference would read.  The container type is kept too.
            old, info = artifacts.read_faiss_index(self.store.index_path)
            vecs = np.asarray(vecs, dtype=np.float32)
            full = np.concatenate([old, vecs]) if old.size else vecs
            self.store.index.add(vecs)
            self.store.chunks.extend(new_chunks)
            self.store.index_path.parent.mkdir(parents=True, exist_ok=True)
            if info.get("hnsw") is not None:
                M = int(getattr(self.cfg.retrieval, "hnsw_m", 64))
                artifacts.write_faiss_hnsw_flat(self.store.index_path, full, M, int(getattr(self.cfg.retrieval, "hnsw_ef_construction", 400)),
                                                int(getattr(self.cfg.retrieval, "hnsw_ef_search", 512)), neighbors=_knn_links(full, 2 * M))
            else:
                artifacts.write_faiss_flat(self.store.index_path, full)
            self.store._index_mtime = self.store.index_path.stat().st_mtime
            self.store._meta_mtime = self.store.meta_path.stat().st_mtime
        return len(new_chunks)

    def add_jsonl(self, path) -> int:
        return self.add_chunks(_load_jsonl_chunks(path))


class IncrementalBM25Builder:
    """builders/incremental_bm25_builder.py:19-82: BM25 statistics are global (idf, avgdl), so an increment
    re-derives the Okapi state over existing + new chunks and replaces bm25.pkl atomically.  Chunks whose id
    is already present are skipped; an unreadable pickle counts as empty, like the reference (:31-41)."""

    def __init__(self, cfg, tokenizer=None):
        self.cfg = cfg
        self.tokenizer = tokenizer

    def _load_existing(self) -> List[LawChunk]:
        path = Path(self.cfg.retrieval.bm25_index_file)
        if not path.exists():
            return []
        try:
            payload = artifacts.read_bm25_pickle(path)
            return [LawChunk.model_validate(c) for c in (payload.get("chunks") or [])]
        except Exception:
            return []

    def add_chunks(self, chunks: Sequence[LawChunk]) -> int:
        if not chunks:
            return 0
        existing = self._load_existing()
        exist_ids = {c.id for c in existing}
        new_chunks = [c for c in chunks if c.id not in exist_ids]
        if not new_chunks:
            return 0
        all_chunks = existing + list(new_chunks)
        tok = self.tokenizer or encoders.default_query_tokenizer()     # the reference re-tokenises with jieba.cut (:70)
        state = artifacts.okapi_state_from_tokens([tok(c.text) for c in all_chunks])
        path = Path(self.cfg.retrieval.bm25_index_file)
        path.parent.mkdir(parents=True, exist_ok=True)
        artifacts.write_bm25_pickle(path, state, all_chunks)             # tmp file + os.replace
        return len(new_chunks)

    def add_jsonl(self, path) -> int:
        return self.add_chunks(_load_jsonl_chunks(path))
