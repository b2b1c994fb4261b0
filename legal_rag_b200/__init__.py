"""legal_rag_b200 -- B200-native engine for Legal-RAG's hybrid-retrieval hot path.

Layers (bottom up):
  csrc/ + include/lrag.h   hand-written sm_100a kernels behind a C ABI (liblrag.so)
  _native                  ctypes binding of that ABI
  engine                   batched tensor API (dense / BM25 / MaxSim / fusion / merge, multi-GPU shards)
  retrieval                the reference's retriever classes (DenseRetriever, BM25Retriever,
                           ColBERTRetriever, HybridRetriever, VectorStore) on top of the engine
"""
__version__ = "0.1.0"
