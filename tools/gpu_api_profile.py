#!/usr/bin/env python
"""Where the host time of HybridRetriever.search / search_batch goes (UCC corpus, hashing stand-in encoders): cProfile of the
steady-state calls only."""
import cProfile, pstats, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from legal_rag_b200.config import AppConfig
from legal_rag_b200.retrieval import HybridRetriever, builders, encoders
from legal_rag_b200.schemas import LawChunk
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(ROOT, "tests", "golden", "ucc_corpus.npz"))
lens, flat, vocab, ids = z["doc_len"], z["tokens"].astype(np.int64), z["vocab"], z["ids"]
off = np.concatenate([[0], np.cumsum(lens)])
texts = [" ".join(vocab[flat[off[i]:off[i + 1]]]) for i in range(len(lens))]
chunks = [LawChunk(id=str(ids[i]), law_name="UCC", article_no=str(i), article_id=str(i), text=texts[i], lang="en") for i in range(len(texts))]
rng = np.random.default_rng(44)
questions = []
for _ in range(256):
    t = texts[int(rng.integers(0, len(texts)))].split()
    L = int(rng.integers(3, 9)); st = int(rng.integers(0, max(1, len(t) - L)))
    questions.append(" ".join(t[st:st + L]))
with tempfile.TemporaryDirectory() as root:
    cfg = AppConfig(); cfg.device = "cuda:0"
    r = cfg.retrieval
    r.faiss_index_file, r.faiss_meta_file = os.path.join(root, "faiss", "faiss.index"), os.path.join(root, "faiss", "faiss_meta.jsonl")
    r.bm25_index_file = os.path.join(root, "bm25.pkl")
    r.enable_colbert = False
    r.top_k, r.min_final_score, r.fusion_method = 100, 0.0, "weighted_sum"
    enc = encoders.HashingDenseEncoder(768)
    encoders.register_dense_encoder(lambda name, dev: enc)
    builders.build_faiss_index(cfg, chunks); builders.build_bm25_index(cfg, chunks)
    hr = HybridRetriever(cfg)
    for q in questions[:16]:
        hr.search(q, None, 100)
    hr.search_batch(questions, 100)
    for rep in range(2):
        t0 = time.perf_counter()
        for q in questions[:128]:
            hr.search(q, None, 100)
        print("search: %.1f us per query" % (1e6 * (time.perf_counter() - t0) / 128))
        t0 = time.perf_counter()
        hr.search_batch(questions, 100)
        print("search_batch: %.0f queries/s" % (256 / (time.perf_counter() - t0)))
    pr = cProfile.Profile(); pr.enable()
    for q in questions[:128]:
        hr.search(q, None, 100)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    pr = cProfile.Profile(); pr.enable()
    hr.search_batch(questions, 100)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
