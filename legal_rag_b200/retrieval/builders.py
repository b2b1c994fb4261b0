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
    it unchanged: vector_store.py:112-117).  flat=False wraps it in an HNSW container with an empty graph."""
    rcfg = cfg.retrieval
    enc = encoder or encoders.make_dense_encoder(str(rcfg.embedding_model), "cpu")
    X = np.asarray(enc.encode([c.text for c in chunks], batch_size=64, max_length=512), dtype=np.float32)
    Path(rcfg.faiss_index_file).parent.mkdir(parents=True, exist_ok=True)
    Path(rcfg.faiss_meta_file).parent.mkdir(parents=True, exist_ok=True)
    if flat:
        artifacts.write_faiss_flat(rcfg.faiss_index_file, X)
    else:
        artifacts.write_faiss_hnsw_flat(rcfg.faiss_index_file, X, int(getattr(rcfg, "hnsw_m", 64)),
                                        int(getattr(rcfg, "hnsw_ef_construction", 400)), int(getattr(rcfg, "hnsw_ef_search", 512)))
    artifacts.write_meta_jsonl(rcfg.faiss_meta_file, chunks)


def build_bm25_index(cfg, chunks: Sequence[LawChunk], tokenizer=None) -> None:
    """builders/bm25_builder.py:22-53: English chunks use the lower-casing regex, others jieba."""
    rcfg = cfg.retrieval
    lang = (getattr(chunks[0], "lang", None) or "zh").strip().lower() if chunks else "zh"
    if tokenizer is None:
        tokenizer = encoders.tokenize_en if lang == "en" else encoders.default_query_tokenizer()
    state = artifacts.okapi_state_from_tokens([tokenizer(c.text) for c in chunks])
    Path(rcfg.bm25_index_file).parent.mkdir(parents=True, exist_ok=True)
    artifacts.write_bm25_pickle(rcfg.bm25_index_file, state, chunks)


class IncrementalDenseBuilder:
    """builders/incremental_dense_builder.py:31-78: append new chunks to the live index; the meta file is
    appended BEFORE the index file is rewritten, so a reader never sees an index row without its chunk."""

    def __init__(self, cfg, store: Optional[VectorStore] = None):
        self.cfg = cfg
        self.store = store or VectorStore.from_config(cfg)

    def add_chunks(self, chunks: Sequence[LawChunk]) -> int:
        if not chunks:
            return 0
        self.store.load()
        vecs = self.store._embed([c.text for c in chunks], is_query=False)
        with open(self.store.meta_path, "a", encoding="utf-8") as f:
            for c in chunks:
                f.write(json.dumps(c.model_dump(), ensure_ascii=False) + "\n")
        self.store.index.add(vecs)
        self.store.chunks.extend(chunks)
        artifacts.write_faiss_flat(self.store.index_path, self.store.index.reconstruct_n())
        self.store._index_mtime = self.store.index_path.stat().st_mtime
        self.store._meta_mtime = self.store.meta_path.stat().st_mtime
        return len(chunks)

    def add_jsonl(self, path) -> int:
        chunks: List[LawChunk] = []
        with open(path, "r", encoding="utf-8") as f:
            for line in f:
                if line.strip():
                    chunks.append(LawChunk.model_validate(json.loads(line)))
        return self.add_chunks(chunks)
