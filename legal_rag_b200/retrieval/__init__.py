"""The reference's retriever classes on top of the B200 engine (same names, signatures, error contracts)."""
from .bm25_retriever import BM25Retriever
from .colbert_retriever import ColBERTRetriever, ColbertRetriever, build_token_store
from .dense_retriever import DenseRetriever
from .graph_retriever import GraphRetriever
from .hybrid_retriever import HybridRetriever
from .vector_store import GpuFlatIndex, VectorStore

__all__ = ["BM25Retriever", "ColBERTRetriever", "ColbertRetriever", "DenseRetriever", "HybridRetriever", "VectorStore",
           "GpuFlatIndex", "GraphRetriever", "build_token_store"]
