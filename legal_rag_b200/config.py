"""The knobs the hot path reads, with the reference's defaults (legalrag/config.py:54-129).  Any object
exposing the same attribute names works in their place (the reference's own AppConfig included)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional


@dataclass
class RetrievalConfig:
    faiss_index_file: str = "index/faiss/faiss.index"
    faiss_meta_file: str = "index/faiss/faiss_meta.jsonl"
    embedding_model: str = "BAAI/bge-base-zh-v1.5"
    hnsw_m: int = 64
    hnsw_ef_construction: int = 400
    hnsw_ef_search: int = 512
    bm25_index_file: str = "index/bm25.pkl"
    enable_graph: bool = True
    graph_seed_k: int = 30
    top_k: int = 10
    bm25_weight: float = 0.4
    dense_weight: float = 0.6
    min_final_score: float = 0.2
    enable_colbert: bool = True
    colbert_index_path: str = ""
    colbert_meta_file: str = "index/colbert/colbert_meta.jsonl"
    colbert_weight: float = 0.35
    colbert_index_name: str = "law"
    colbert_model_name: str = "jinaai/jina-colbert-v2"
    colbert_experiment: str = "experiment"
    colbert_nranks: int = 1
    colbert_nbits: int = 4
    colbert_doc_maxlen: int = 220
    enable_rerank: bool = True
    rerank_top_n: int = 30
    rrf_alpha: float = 0.5
    rerank_beta: float = 0.35
    fusion_method: str = "rrf_norm_blend"
    rrf_k: int = 60
    rrf_blend_alpha: float = 0.6     # never read by the reference either


@dataclass
class AppConfig:
    retrieval: RetrievalConfig = field(default_factory=RetrievalConfig)
    device: Optional[str] = None     # engine extension: CUDA device of this process (default: current)
