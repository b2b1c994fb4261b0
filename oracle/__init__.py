"""CPU oracle for the Legal-RAG hybrid-retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``legal_rag_b200/`` may import this
package.  The only legal importers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and
there only as the checker or as the timed CPU baseline, never as the product.

Pinning status ("Oracle"):

* fusion  -- PINNED: ``tests/golden/fuse_golden.json`` was produced by
  executing the reference's own ``HybridRetriever._fuse`` (unmodified source at
  /root/reference/legalrag/retrieval/hybrid_retriever.py:389-551) in this
  container via ``oracle/make_golden.py``; ``oracle/fuse.py`` is checked against
  it in ``tests/test_oracle.py``.
* graph-expansion scoring -- PINNED: ``tests/golden/graph_golden.json`` was produced by executing the reference's
  own ``GraphRetriever.search`` (graph_retriever.py:85-219) on fake collaborators; ``oracle/graph.py`` is checked
  against it.
* BM25 -- additionally checked against the one known answer rank_bm25 publishes (its README example,
  ``array([0., 0.93729472, 0.])``), which the transcription reproduces to 5e-9; everything else about it is, like
* dense / BM25 / MaxSim -- parity unpinned: the arithmetic lives in the
  third-party wheels faiss-cpu 1.13.2, rank-bm25 0.2.2 and colbert-ai 0.2.22,
  none of which is vendored under /root/reference or installed here, and the
  reference's own tests pin no number at that boundary.
  The restatements follow the published algorithms and the reference's call
  sites (cited per function); hand-computed known-answer vectors live in
  ``tests/golden/``.
"""
