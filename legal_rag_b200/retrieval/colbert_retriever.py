"""ColBERTRetriever (replaces legalrag/retrieval/colbert_retriever.py:29-183).

The reference delegates to Stanford ColBERT's PLAID Searcher (centroid pruning + 4-bit residuals).  Here the
channel is an EXACT MaxSim scan of a bf16 token store resident in HBM (liblrag maxsim kernel): at the
reference's corpus sizes (10^3 docs) a full scan is microseconds, and exact MaxSim is what PLAID
approximates.  The token store is this engine's own artifact, written next to the reference's
colbert_meta.jsonl by `build_token_store`; when it is absent but the directory holds a PLAID index written by
colbert's Indexer, that index is decompressed into the same store at load time (retrieval/plaid.py)."""
from __future__ import annotations

from pathlib import Path
from typing import ClassVar, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import engine
from ..schemas import LawChunk
from . import artifacts, encoders, plaid

TOKEN_STORE_FILE = "lrag_token_store.npz"
QUERY_MAXLEN = 32
DIM = 128


def token_store_path(rcfg) -> Path:
    return (Path(str(rcfg.colbert_index_path)) / str(rcfg.colbert_experiment) / "indexes" / str(rcfg.colbert_index_name)
            / TOKEN_STORE_FILE)


def _bf16_bits(x: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).view(torch.int16).numpy()


def build_token_store(cfg, chunks: Sequence[LawChunk], encoder=None) -> Path:
    """Offline counterpart of builders/colbert_builder.py:55-136 for this engine: encode every chunk to
    unit 128-d token vectors, pad to a common length (32, 64, 128 or 256 rows: lengths that tile the 256-row
    block of the batched full-corpus scan) and save bf16 bits + lengths."""
    rcfg = cfg.retrieval
    enc = encoder or encoders.make_token_encoder(str(rcfg.colbert_model_name), "cpu",
                                                 doc_maxlen=int(getattr(rcfg, "colbert_doc_maxlen", 220)))
    mats = [np.asarray(enc.encode_doc((c.text or "").strip()), dtype=np.float32)[: int(getattr(rcfg, "colbert_doc_maxlen", 220))]
            for c in chunks]
    longest = max(m.shape[0] for m in mats)
    if longest > 256:
        raise ValueError(f"documents of {longest} tokens exceed the 256-token tile of the MaxSim kernel")
    Ld = next(n for n in (32, 64, 128, 256) if n >= longest)
    toks = np.zeros((len(mats), Ld, DIM), dtype=np.float32)
    doclen = np.zeros(len(mats), dtype=np.int32)
    for i, m in enumerate(mats):
        toks[i, : m.shape[0]] = m
        doclen[i] = m.shape[0]
    path = token_store_path(rcfg)
    path.parent.mkdir(parents=True, exist_ok=True)
    np.savez(path, tokens_bf16=_bf16_bits(toks), doclen=doclen)
    Path(rcfg.colbert_meta_file).parent.mkdir(parents=True, exist_ok=True)
    artifacts.write_colbert_meta(rcfg.colbert_meta_file, chunks)
    return path


class ColBERTRetriever:
    _instances_by_key: ClassVar[Dict[Tuple, "ColBERTRetriever"]] = {}

    def __init__(self, cfg, encoder=None):
        self.cfg = cfg
        rcfg = cfg.retrieval
        self.enabled: bool = bool(getattr(rcfg, "enable_colbert", False))
        self.index_path = Path(str(getattr(rcfg, "colbert_index_path")))
        self.index_name = str(getattr(rcfg, "colbert_index_name"))
        self.model_name: Optional[str] = getattr(rcfg, "colbert_model_name", "colbert-ir/colbertv2.0")
        self.meta_file = Path(str(getattr(rcfg, "colbert_meta_file")))
        self.experiment = str(getattr(rcfg, "colbert_experiment"))
        self.nranks = int(getattr(rcfg, "colbert_nranks", 1))
        self.device = torch.device(getattr(cfg, "device", None) or "cuda")
        self._pid2chunk: Dict[int, LawChunk] = {}
        self._meta_mtime: Optional[float] = None
        self._store_mtime: Optional[float] = None
        self._tokens: Optional[torch.Tensor] = None      # [Nd, Ld, 128] bf16 on the device
        self._doclen: Optional[torch.Tensor] = None
        self._encoder = encoder
        if not self.enabled:
            return
        self._load_meta_and_collection()
        self._load_token_store()
        if self._encoder is None:
            self._encoder = encoders.make_token_encoder(str(self.model_name), self.device, query_maxlen=QUERY_MAXLEN,
                                                        doc_maxlen=int(getattr(rcfg, "colbert_doc_maxlen", 220)))

    @classmethod
    def from_config(cls, cfg) -> "ColBERTRetriever":
        rcfg = cfg.retrieval
        key = (str(getattr(rcfg, "colbert_index_path")), str(getattr(rcfg, "colbert_index_name")),
               str(getattr(rcfg, "colbert_model_name", "colbert-ir/colbertv2.0")), str(getattr(rcfg, "colbert_meta_file")),
               str(getattr(rcfg, "colbert_experiment")), int(getattr(rcfg, "colbert_nranks", 1)))
        inst = cls._instances_by_key.get(key)
        if inst is None:
            inst = cls._instances_by_key[key] = cls(cfg)
        return inst

    def _load_meta_and_collection(self) -> None:
        if not self.meta_file.exists():
            raise RuntimeError(f"ColBERT meta file not found: {self.meta_file}. Run build_colbert_index() first.")
        meta_mtime = self.meta_file.stat().st_mtime
        if self._meta_mtime is not None and self._meta_mtime == meta_mtime:
            return
        pid2chunk = artifacts.read_colbert_meta(self.meta_file)
        if not pid2chunk:
            raise RuntimeError(f"ColBERT meta file is empty: {self.meta_file}")
        self._pid2chunk, self._meta_mtime = pid2chunk, meta_mtime

    def _load_token_store(self) -> None:
        path = token_store_path(self.cfg.retrieval)
        if not path.exists():
            if plaid.is_plaid_dir(path.parent):
                # an index written by colbert's Indexer (builders/colbert_builder.py:109-134): decompress it once
                mtime = (path.parent / "metadata.json").stat().st_mtime
                if self._store_mtime == mtime and self._tokens is not None:
                    return
                toks, doclen = plaid.read_plaid_index(path.parent, self.device)
                self._tokens, self._doclen, self._store_mtime = toks, doclen, mtime
                return
            raise RuntimeError(f"ColBERT token store not found: {path}. Run legal_rag_b200.retrieval.build_token_store() first.")
        mtime = path.stat().st_mtime
        if self._store_mtime == mtime and self._tokens is not None:
            return
        z = np.load(path)
        toks = torch.from_numpy(z["tokens_bf16"]).view(torch.bfloat16).to(self.device)
        doclen = torch.from_numpy(z["doclen"].astype(np.int32)).to(self.device)
        self._tokens, self._doclen, self._store_mtime = toks, doclen, mtime

    def _encode_queries(self, queries: Sequence[str]) -> torch.Tensor:
        if hasattr(self._encoder, "encode_queries_device"):
            # checkpoint-backed encoder on the index's device: text -> kernel operand without a host round trip
            Q = self._encoder.encode_queries_device(list(queries))[:, :QUERY_MAXLEN]
            return Q.to(self.device).to(torch.bfloat16).contiguous()
        mats = [np.asarray(self._encoder.encode_query(q), dtype=np.float32)[:QUERY_MAXLEN] for q in queries]
        Lq = max(m.shape[0] for m in mats)
        Q = np.zeros((len(mats), Lq, DIM), dtype=np.float32)     # zero rows add max_j <0, d_j> = 0 to every doc
        for i, m in enumerate(mats):
            Q[i, : m.shape[0]] = m
        return torch.from_numpy(Q).to(self.device).to(torch.bfloat16)

    def search_device(self, queries: Sequence[str], top_k: int):
        """-> (scores [nq, k], pids [nq, k]) on the device: exact MaxSim over every document."""
        self._load_token_store()
        Nd = self._tokens.shape[0]
        k = max(1, min(int(top_k), Nd, engine.LRAG_MAX_K))
        Q = self._encode_queries(queries)
        if Q.shape[0] > 1 and engine.maxsim_scan_supported(int(self._tokens.shape[1])):
            # a batch shares one pass over the token store (tensor-bound contraction) instead of one gather per query
            return engine.maxsim_scan_topk(self._tokens, self._doclen, Q, k)
        cand = torch.arange(Nd, device=self.device, dtype=torch.int64).unsqueeze(0).expand(Q.shape[0], Nd).contiguous()
        return engine.maxsim_rerank(self._tokens, self._doclen, Q, cand, k)

    def rerank_device(self, queries: Sequence[str], cand: torch.Tensor, top_k: int):
        """Candidate mode (north star: top-1000 -> top-100): cand [nq, C] int64 pids, -1 = skip."""
        self._load_token_store()
        k = max(1, min(int(top_k), cand.shape[1], engine.LRAG_MAX_K))
        return engine.maxsim_rerank(self._tokens, self._doclen, self._encode_queries(queries), cand.to(self.device), k)

    def search(self, query: str, top_k: int = 5) -> List[Tuple[LawChunk, float]]:
        if not self.enabled:
            return []
        if self._tokens is None:
            raise RuntimeError("ColBERT token store is not initialized.")
        self._load_meta_and_collection()
        query = (query or "").strip()
        if not query:
            return []
        try:
            s, i = self.search_device([query], top_k)
            s, i = s[0].tolist(), i[0].tolist()
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()          # colbert_retriever.py:153-173: an OOM degrades to no hits
            return []
        out: List[Tuple[LawChunk, float]] = []
        for pid, score in zip(i, s):
            chunk = self._pid2chunk.get(int(pid))
            if chunk is not None and pid >= 0:
                out.append((chunk, float(score)))
        return out


ColbertRetriever = ColBERTRetriever     # the north star's spelling
