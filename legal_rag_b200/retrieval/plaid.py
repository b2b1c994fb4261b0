"""Reader (and, for tests, writer) of Stanford ColBERT's PLAID index directories -- SURVEY 8f rank 3.

The reference builds its ColBERT channel with `colbert.Indexer` (legalrag/retrieval/builders/colbert_builder.py:
109-134, nbits = `colbert_nbits` = 4, config.py:107) and searches it with `colbert.Searcher`
(colbert_retriever.py:119-152).  This module turns such a directory into the dense bf16 token store the MaxSim
kernels scan, so an existing index can be served without re-encoding the corpus: every token embedding is
rebuilt as  normalize(centroid[code] + bucket_weight[residual bucket])  -- what `ResidualCodec.decompress`
computes before scoring -- and laid out as [Nd, Ld, 128] with per-document lengths.

PARITY UNPINNED: colbert-ai (0.2.22 resolved in notebooks/02_LegalRAG_Pipeline.ipynb:420) is neither vendored
under /root/reference nor installable here, and the reference ships no index directory.  The layout below is
restated from the library's published indexer / codec:

  metadata.json            {"num_chunks", "num_embeddings", "num_partitions", "avg_doclen", "config": {"nbits", "dim", ...}}
  centroids.pt             half [num_partitions, dim]
  buckets.pt               (bucket_cutoffs [2^nbits - 1], bucket_weights [2^nbits])
  avg_residual.pt          scalar / [dim]                (not needed for decompression)
  {c}.codes.pt             int32 [n_c]                   centroid id of every embedding of chunk c
  {c}.residuals.pt         uint8 [n_c, dim * nbits / 8]  bucket indices, bit-packed
  doclens.{c}.json         [int]                         embeddings per passage of chunk c
  {c}.metadata.json        {"passage_offset", "num_passages", "num_embeddings", "embedding_offset"}
  ivf.pid.pt               (ivf, ivf_lengths)            candidate generation only: ignored (the scan is exact)

Bit packing (`ResidualCodec.binarize`): a bucket index v is expanded to its nbits bits LEAST significant first and the
bit stream is packed with numpy.packbits (most significant bit of a byte first).  For nbits = 4 a byte therefore holds
two indices, the first in the high nibble, each nibble bit-reversed.  The writer here follows the same recipe, so the
round trip in tests/test_host_logic.py checks this module against itself only; verify against a real index before relying
on it.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import List, Tuple

import numpy as np
import torch

DIM = 128


def _bitrev(v: int, nbits: int) -> int:
    return int(format(v, f"0{nbits}b")[::-1], 2)


def _unpack_lut(nbits: int) -> torch.Tensor:
    """[256, 8 / nbits] uint8: the bucket indices held by each possible byte, in embedding-dimension order."""
    per = 8 // nbits
    mask = (1 << nbits) - 1
    lut = np.zeros((256, per), dtype=np.uint8)
    for b in range(256):
        for j in range(per):
            lut[b, j] = _bitrev((b >> (8 - nbits * (j + 1))) & mask, nbits)
    return torch.from_numpy(lut)


def is_plaid_dir(path) -> bool:
    p = Path(path)
    return (p / "metadata.json").exists() and (p / "centroids.pt").exists() and (p / "buckets.pt").exists()


def read_plaid_index(index_dir, device="cpu", pad_to: Tuple[int, ...] = (32, 64, 128, 256)) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (tokens bf16 [Nd, Ld, dim] zero-padded, doclen int32 [Nd]); pid = row, in index order."""
    d = Path(index_dir)
    meta = json.loads((d / "metadata.json").read_text())
    cfg = meta.get("config", {})
    nbits = int(cfg.get("nbits", 4))
    dim = int(cfg.get("dim", DIM))
    if 8 % nbits:
        raise ValueError(f"PLAID index with nbits={nbits}: only 1, 2, 4 or 8 bits per dimension are packed byte-aligned")
    centroids = torch.load(d / "centroids.pt", map_location="cpu").float().to(device)
    buckets = torch.load(d / "buckets.pt", map_location="cpu")
    weights = buckets[1].float().to(device)
    lut = _unpack_lut(nbits).to(device)
    doclens: List[int] = []
    embs = []
    for c in range(int(meta["num_chunks"])):
        codes = torch.load(d / f"{c}.codes.pt", map_location="cpu").long().to(device)
        packed = torch.load(d / f"{c}.residuals.pt", map_location="cpu").to(device)
        idx = lut[packed.long()].reshape(packed.shape[0], -1)[:, :dim]          # [n, dim] bucket index per dimension
        e = centroids[codes] + weights[idx.long()]
        embs.append(torch.nn.functional.normalize(e, p=2, dim=-1).to(torch.bfloat16))
        doclens += [int(x) for x in json.loads((d / f"doclens.{c}.json").read_text())]
    flat = torch.cat(embs) if embs else torch.zeros((0, dim), dtype=torch.bfloat16, device=device)
    if sum(doclens) != flat.shape[0]:
        raise ValueError(f"PLAID index {d}: doclens sum to {sum(doclens)} but {flat.shape[0]} embeddings are stored")
    longest = max(doclens) if doclens else 1
    fits = [n for n in pad_to if n >= longest]
    if not fits:
        raise ValueError(f"PLAID index {d}: a passage has {longest} embeddings, the MaxSim tile holds {max(pad_to)}")
    Ld = fits[0]
    dl = torch.tensor(doclens, dtype=torch.int64, device=device)
    tokens = torch.zeros((len(doclens), Ld, dim), dtype=torch.bfloat16, device=device)
    if flat.shape[0]:
        doc_of = torch.repeat_interleave(torch.arange(len(doclens), device=device), dl)
        pos = torch.arange(flat.shape[0], device=device) - torch.repeat_interleave(torch.cumsum(dl, 0) - dl, dl)
        tokens[doc_of, pos] = flat
    return tokens, dl.to(torch.int32)


def write_plaid_index(index_dir, doc_embs: List[np.ndarray], n_centroids: int = 16, nbits: int = 4, chunk_docs: int = 0,
                      seed: int = 0) -> None:
    """Test helper: compress unit-norm token embeddings the way colbert's indexer does (nearest centroid by inner
    product, per-dimension residual bucketised at the quantile cutoffs, bits packed LSB-first with numpy.packbits)."""
    d = Path(index_dir)
    d.mkdir(parents=True, exist_ok=True)
    allv = np.concatenate(doc_embs).astype(np.float32)
    dim = allv.shape[1]
    rng = np.random.default_rng(seed)
    cent = allv[rng.choice(len(allv), size=min(n_centroids, len(allv)), replace=False)].copy()
    codes_all = np.argmax(allv @ cent.T, axis=1)
    res = allv - cent[codes_all]
    nb = 1 << nbits
    cutoffs = np.quantile(res, np.arange(1, nb) / nb).astype(np.float32)
    weights = np.quantile(res, (np.arange(nb) + 0.5) / nb).astype(np.float32)
    torch.save(torch.from_numpy(cent).half(), d / "centroids.pt")
    torch.save((torch.from_numpy(cutoffs), torch.from_numpy(weights)), d / "buckets.pt")
    torch.save(torch.from_numpy(np.abs(res).mean(axis=0)).half(), d / "avg_residual.pt")
    chunk_docs = chunk_docs or len(doc_embs)
    chunks = [doc_embs[i:i + chunk_docs] for i in range(0, len(doc_embs), chunk_docs)]
    off_e = off_p = 0
    for c, docs in enumerate(chunks):
        v = np.concatenate(docs).astype(np.float32)
        codes = np.argmax(v @ cent.T, axis=1).astype(np.int32)
        bucket = np.searchsorted(cutoffs, v - cent[codes], side="left").astype(np.uint8)       # torch.bucketize(right=False)
        bits = (bucket[:, :, None] >> np.arange(nbits, dtype=np.uint8)) & 1                  # least significant bit first
        packed = np.packbits(bits.reshape(-1)).reshape(len(v), dim * nbits // 8)
        torch.save(torch.from_numpy(codes), d / f"{c}.codes.pt")
        torch.save(torch.from_numpy(packed), d / f"{c}.residuals.pt")
        (d / f"doclens.{c}.json").write_text(json.dumps([int(len(x)) for x in docs]))
        (d / f"{c}.metadata.json").write_text(json.dumps({"passage_offset": off_p, "num_passages": len(docs),
                                                         "num_embeddings": int(len(v)), "embedding_offset": off_e}))
        off_e += len(v)
        off_p += len(docs)
    (d / "metadata.json").write_text(json.dumps({"config": {"nbits": nbits, "dim": dim}, "num_chunks": len(chunks),
                                                "num_partitions": int(len(cent)), "num_embeddings": int(off_e),
                                                "avg_doclen": off_e / max(1, off_p)}))
