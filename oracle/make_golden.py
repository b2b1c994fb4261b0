"""Generate tests/golden/* by EXECUTING the reference (only works where
/root/reference exists; the outputs are committed so the GPU box never needs it).

  python -m oracle.make_golden            # writes tests/golden/fuse_golden.json, ucc_corpus.npz, graph_golden.json

1. fuse_golden.json -- the reference's own ``HybridRetriever._fuse`` /
   ``_minmax`` / ``_rrf_with_breakdown`` / ``_dedup_keep_best``
   (legalrag/retrieval/hybrid_retriever.py) run unmodified; third-party modules that are
   not installed here (faiss, jieba, rank_bm25, colbert, FlagEmbedding,
   sentence_transformers) are replaced by empty stubs in ``sys.modules`` -- none of them
   is touched by the fusion code.
2. ucc_corpus.npz -- BASELINE.json config 1's corpus: the reference's
   ``scripts.preprocess_law.parse_by_lines`` over data/raw/ucc/*.txt, id-dedup as
   ``legalrag/retrieval/corpus_loader.py:35-38``, tokenised with the reference's
   ``_tokenize_en`` regex (builders/bm25_builder.py:18-19); stored as term ids.
"""
from __future__ import annotations

import json
import random
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _stub_modules() -> None:
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Dummy:  # any attribute / call is accepted
        def __init__(self, *a, **k):
            pass

    mod("faiss")
    mod("jieba", cut=lambda s: s.split())
    mod("rank_bm25", BM25Okapi=_Dummy)
    mod("colbert", Searcher=_Dummy, Indexer=_Dummy)
    mod("colbert.infra", Run=_Dummy, RunConfig=_Dummy, ColBERTConfig=_Dummy)
    mod("FlagEmbedding", FlagModel=_Dummy)
    mod("sentence_transformers", CrossEncoder=_Dummy, SentenceTransformer=_Dummy)


def _load_reference():
    _stub_modules()
    sys.path.insert(0, str(REF))
    from legalrag.config import RetrievalConfig
    from legalrag.retrieval import hybrid_retriever as hr
    from legalrag.schemas import LawChunk, RetrievalHit
    return hr, RetrievalConfig, LawChunk, RetrievalHit


def _chunk(LawChunk, i):
    return LawChunk(id=f"d{i}", law_name="L", article_no=str(i), article_id=str(i), text=f"t{i}")


def _hits(LawChunk, RetrievalHit, pairs, ch):
    return [RetrievalHit(chunk=_chunk(LawChunk, i), score=float(s), rank=r + 1,
                         score_breakdown={"channel": [ch], f"{ch}_raw": float(s)})
            for r, (i, s) in enumerate(pairs)]


def _run_fuse(hr, RetrievalConfig, LawChunk, RetrievalHit, case, method, overrides=None):
    rcfg = RetrievalConfig()
    rcfg.fusion_method = method
    for k, v in (overrides or {}).items():
        setattr(rcfg, k, v)
    obj = object.__new__(hr.HybridRetriever)
    obj.cfg = SimpleNamespace(retrieval=rcfg)
    out = obj._fuse(dense_hits=_hits(LawChunk, RetrievalHit, case["dense"], "dense"),
                    bm25_hits=_hits(LawChunk, RetrievalHit, case["bm25"], "bm25"),
                    colbert_hits=_hits(LawChunk, RetrievalHit, case["colbert"], "colbert"))
    rows = []
    for h in out:
        sb = h.score_breakdown
        rows.append({"id": int(h.chunk.id[1:]), "score": h.score, "rank": h.rank,
                     "rrf_norm": sb["rrf_norm"], "weighted_sum": sb["weighted_sum"],
                     "dense_norm": sb["dense_norm"], "bm25_norm": sb["bm25_norm"],
                     "colbert_norm": sb["colbert_norm"], "channel": sb["channel"],
                     "channel_contrib": sb["channel_contrib"]})
    return {"method": method, "rrf_k": rcfg.rrf_k, "alpha": rcfg.rrf_alpha,
            "w_dense": rcfg.dense_weight, "w_bm25": rcfg.bm25_weight, "w_colbert": rcfg.colbert_weight,
            "out": rows}


def make_fuse_golden() -> None:
    hr, RetrievalConfig, LawChunk, RetrievalHit = _load_reference()
    cases = {
        # SURVEY.md Appendix A
        "A1": {"dense": [(1, .65), (2, .53), (3, .51)], "bm25": [(1, 37.36), (4, 8.09), (2, 7.5)],
               "colbert": [(1, 22.08), (3, 19.97), (5, 19.9)]},
        "E1": {"dense": [(1, .9), (2, .5), (3, .1)], "bm25": [(2, 10), (4, 4), (5, 0), (6, 0)], "colbert": []},
        "E2": {"dense": [(1, .5), (2, .5)], "bm25": [(2, 3), (1, 1)], "colbert": []},
        "E3": {"dense": [(1, .9)], "bm25": [(2, 3), (1, 1)], "colbert": []},
        "E4": {"dense": [], "bm25": [], "colbert": []},
        "E5": {"dense": [(1, -0.1), (2, -0.5)], "bm25": [], "colbert": []},
    }
    rng = random.Random(42)   # the reference's own seed convention (scripts/generate_synthetic_data.py:36)
    for n, (kd, kb, kc, universe) in enumerate([(10, 10, 10, 25), (100, 100, 100, 180), (100, 100, 0, 150),
                                                (37, 100, 5, 120), (100, 3, 100, 400), (1, 1, 1, 2)]):
        def lst(kk, lo, hi):
            ids = rng.sample(range(universe), min(kk, universe))
            sc = sorted((rng.uniform(lo, hi) for _ in ids), reverse=True)
            return [(i, s) for i, s in zip(ids, sc)]
        cases[f"R{n}"] = {"dense": lst(kd, 0.2, 0.9), "bm25": lst(kb, 0.0, 40.0), "colbert": lst(kc, 10.0, 30.0)}

    golden = {"source": "legalrag/retrieval/hybrid_retriever.py:_fuse executed by oracle/make_golden.py",
              "cases": {}}
    for name, case in cases.items():
        runs = [_run_fuse(hr, RetrievalConfig, LawChunk, RetrievalHit, case, m)
                for m in ("weighted_sum", "rrf", "wrrf", "rrf_norm_blend")]
        # non-default knobs on the random cases
        if name.startswith("R"):
            runs.append(_run_fuse(hr, RetrievalConfig, LawChunk, RetrievalHit, case, "rrf_norm_blend",
                                  {"rrf_k": 10, "rrf_alpha": 0.25, "dense_weight": 0.5,
                                   "bm25_weight": 0.3, "colbert_weight": 0.2}))
        golden["cases"][name] = {"inputs": case, "runs": runs}

    # helper known-answers (Appendix A E6/E7)
    t, _ = hr._rrf_with_breakdown({"dense": ["a", "b"], "bm25": ["b", "c"]}, k=60)
    tw, _ = hr._rrf_with_breakdown({"dense": ["a", "b"], "bm25": ["b", "c"]}, k=60,
                                   weights={"dense": .6, "bm25": .4})
    golden["helpers"] = {
        "rrf_E6": t, "wrrf_E6": tw,
        "minmax": [[[3, 1, 2], hr._minmax([3, 1, 2])], [[1, 1 + 1e-13], hr._minmax([1, 1 + 1e-13])],
                   [[], hr._minmax([])]],
    }
    # E8 dedup
    c1, c2 = _chunk(LawChunk, 1), _chunk(LawChunk, 2)
    hs = [RetrievalHit(chunk=c1, score=0.4, score_breakdown={"channel": ["dense"], "channel_contrib": {"dense": .4}}),
          RetrievalHit(chunk=c1, score=0.7, source="graph", score_breakdown={"channel": "graph"}),
          RetrievalHit(chunk=c2, score=0.5, score_breakdown={"channel": ["bm25"]})]
    dd = hr._dedup_keep_best(hs)
    golden["helpers"]["dedup_E8"] = [{"id": h.chunk.id, "score": h.score, "rank": h.rank, "source": h.source,
                                      "channel": sorted(h.score_breakdown.get("channel", [])),
                                      "channel_contrib": h.score_breakdown.get("channel_contrib")} for h in dd]
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "fuse_golden.json").write_text(json.dumps(golden, separators=(",", ":")))
    print("wrote", OUT / "fuse_golden.json", "cases:", len(golden["cases"]))


def make_ucc_corpus() -> None:
    sys.path.insert(0, str(REF))
    from scripts.preprocess_law import parse_by_lines
    from oracle.bm25 import tokenize_en

    records = []
    for p in sorted((REF / "data" / "raw" / "ucc").glob("*.txt")):
        text = p.read_text(encoding="utf-8", errors="ignore")
        if not text.strip():
            continue
        records += parse_by_lines(text, source=p.name, law_name=p.stem)
    seen, docs = set(), []
    for r in records:                       # corpus_loader.py:35-38 id-dedup
        if r["id"] in seen:
            continue
        seen.add(r["id"])
        docs.append(r)
    toks = [tokenize_en(r["text"]) for r in docs]
    vocab = sorted({w for d in toks for w in d})
    w2i = {w: i for i, w in enumerate(vocab)}
    flat = np.array([w2i[w] for d in toks for w in d], dtype=np.uint16)
    lens = np.array([len(d) for d in toks], dtype=np.int32)
    np.savez_compressed(OUT / "ucc_corpus.npz", tokens=flat, doc_len=lens,
                        vocab=np.array(vocab), ids=np.array([r["id"] for r in docs]))
    print("UCC: records", len(records), "docs", len(docs), "tokens", int(lens.sum()), "V", len(vocab))


def make_graph_golden() -> None:
    """graph_golden.json -- the scoring stage of the reference's GraphRetriever.search
    (legalrag/retrieval/graph_retriever.py:85-219) run unmodified on fake collaborators: a graph whose
    walk() returns a fixed node list, and a store whose _embed() returns fixed vectors per text."""
    _stub_modules()
    sys.path.insert(0, str(REF))
    from legalrag.config import RetrievalConfig
    from legalrag.retrieval import graph_retriever as gr
    from legalrag.schemas import LawChunk

    rng = np.random.default_rng(31)
    d, n = 64, 40
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    qvec = rng.standard_normal(d).astype(np.float32)
    qvec /= np.linalg.norm(qvec)
    chunks = [LawChunk(id=f"c{i}", law_name="L", article_no=str(i), article_id=f"a{i}", text=f"text {i}",
                       lang="en" if i % 5 else "zh") for i in range(n)]
    chunks[7].text = "   "                                        # blank text: dropped (:150)
    text2vec = {c.text: vecs[i] for i, c in enumerate(chunks)}
    rel_pool = [[], ["cite"], ["next"], ["defined_by", "next"], ["neighbor"], ["amend"], ["CITED"], ["unknown_rel"], ["prev", "ref"]]

    class Node(SimpleNamespace):
        pass

    nodes = []
    for j in range(60):
        i = int(rng.integers(0, n + 3))                           # a few article ids are not in the store (:148-150)
        nodes.append(Node(article_id=f"a{i}", graph_depth=int(rng.integers(0, 4)), relations=rel_pool[j % len(rel_pool)],
                          meta={"_edge_conf": float(rng.choice([1.0, 0.9, 0.5, 0.0]))} if j % 3 else {}))
    nodes.append(Node(article_id="", graph_depth=1, relations=[], meta={}))

    class Store:
        def _embed(self, texts, is_query=False):
            if isinstance(texts, str):
                return qvec
            return np.stack([text2vec[t] for t in texts])

    class Graph:
        def walk(self, **kw):
            self.kw = kw
            return nodes

    cases = []
    for lang, top_k, gamma in ((None, 10, 0.7), ("en", 50, 0.7), (None, 5, 1.3)):
        ret = object.__new__(gr.GraphRetriever)
        base = RetrievalConfig()          # has no graph_depth_gamma field: the reference's getattr default (0.7) applies
        rcfg = SimpleNamespace(graph_walk_depths=base.graph_walk_depths, graph_limit=base.graph_limit,
                               graph_rel_types=base.graph_rel_types, graph_min_conf=0.0, graph_depth_gamma=gamma)
        ret.cfg = SimpleNamespace(retrieval=rcfg)
        ret.graph, ret.store = Graph(), Store()
        ret.id2chunk = {c.article_id: c for c in chunks}
        seeds = [SimpleNamespace(chunk=chunks[0]), SimpleNamespace(chunk=chunks[3])]
        hits = ret.search("the question", seeds, lang=lang, top_k=top_k)
        cases.append({"lang": lang, "top_k": top_k, "gamma": gamma,
                      "hits": [{"id": h.chunk.id, "score": h.score, "rank": h.rank, "source": h.source,
                                "breakdown": {k: h.score_breakdown[k] for k in ("semantic", "depth_decay", "relation_weight", "edge_conf", "final", "graph_depth")}}
                               for h in hits]})
    out = {"d": d, "vectors": vecs.tolist(), "qvec": qvec.tolist(),
           "chunks": [c.model_dump() for c in chunks],
           "nodes": [{"article_id": x.article_id, "graph_depth": x.graph_depth, "relations": x.relations, "meta": x.meta} for x in nodes],
           "cases": cases,
           "functions": {"depth_decay": [[dpt, g, gr._depth_decay(dpt, gamma=g)] for dpt in (-1, 0, 1, 2, 5) for g in (0.7, 1.0)],
                         "relation_weight": [[r, gr._relation_weight(r)] for r in rel_pool]}}
    (OUT / "graph_golden.json").write_text(json.dumps(out))
    print("graph golden:", [len(c["hits"]) for c in cases])


if __name__ == "__main__":
    make_ucc_corpus()
    make_fuse_golden()
    make_graph_golden()
