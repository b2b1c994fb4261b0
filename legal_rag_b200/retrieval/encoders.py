"""Text -> vectors / tokens.  The encoders (BGE for the dense channel, a ColBERT checkpoint for late
interaction, jieba for BM25 queries) are NOT part of the hot path this package accelerates (SURVEY 8f ranks
them "next"); they are pluggable here.  The library-backed defaults are used when the reference's
dependencies are installed; tests and benchmarks plug in deterministic stand-ins."""
from __future__ import annotations

import hashlib
import re
from typing import Callable, List, Optional, Sequence

import numpy as np

_TOKEN_RE = re.compile(r"[A-Za-z0-9]+(?:'[A-Za-z0-9]+)?|[一-鿿]|\s+|[^\sA-Za-z0-9一-鿿]")


def tokenize_en(text: str) -> List[str]:
    """The reference's English INDEX tokenizer (builders/bm25_builder.py:18-19)."""
    return re.findall(r"[A-Za-z0-9]+(?:'[A-Za-z0-9]+)?", text.lower())


def default_query_tokenizer() -> Callable[[str], List[str]]:
    """bm25_retriever.py:73 tokenises QUERIES with jieba.cut for both languages.  Without jieba a regex
    splitter keeps its observable behaviour on ASCII text (case preserved; words, whitespace runs and
    punctuation become separate tokens) and falls back to single CJK characters."""
    try:
        import jieba
        return lambda q: list(jieba.cut(q))
    except ImportError:
        return lambda q: _TOKEN_RE.findall(q)


class HashingDenseEncoder:
    """Deterministic stand-in for BGE (tests / offline use): signed feature hashing of word tokens into
    `dim` buckets, L2-normalised float32 like FlagModel output (vector_store.py:131-155)."""

    def __init__(self, dim: int = 768):
        self.dim = dim

    def _one(self, text: str) -> np.ndarray:
        v = np.zeros(self.dim, dtype=np.float32)
        for tok in tokenize_en(text) or [text]:
            h = int.from_bytes(hashlib.blake2b(tok.encode("utf-8"), digest_size=8).digest(), "little")
            v[h % self.dim] += 1.0 if (h >> 40) & 1 else -1.0
        n = float(np.linalg.norm(v))
        return v / n if n > 0 else v

    def encode(self, texts: Sequence[str], **_) -> np.ndarray:
        return np.stack([self._one(t) for t in texts]) if len(texts) else np.zeros((0, self.dim), np.float32)

    encode_queries = encode


class HashingTokenEncoder:
    """Deterministic stand-in for a ColBERT checkpoint: one unit 128-d vector per word token."""

    def __init__(self, dim: int = 128, query_maxlen: int = 32, doc_maxlen: int = 220):
        self.dim, self.query_maxlen, self.doc_maxlen = dim, query_maxlen, doc_maxlen

    def _tok(self, tok: str) -> np.ndarray:
        seed = int.from_bytes(hashlib.blake2b(tok.encode("utf-8"), digest_size=8).digest(), "little")
        v = np.random.default_rng(seed).standard_normal(self.dim).astype(np.float32)
        return v / np.linalg.norm(v)

    def encode_query(self, text: str) -> np.ndarray:
        toks = tokenize_en(text)[: self.query_maxlen] or ["[empty]"]
        return np.stack([self._tok(t) for t in toks])

    def encode_doc(self, text: str) -> np.ndarray:
        toks = tokenize_en(text)[: self.doc_maxlen] or ["[empty]"]
        return np.stack([self._tok(t) for t in toks])


BGE_QUERY_INSTRUCTION = "为这个法律问题生成表示以用于检索相关法律条文："        # vector_store.py:72 (used for both languages)


class TransformersBgeEncoder:
    """BGE-style dense encoder on the same GPU as the index, for installs that have the checkpoint on
    disk and `transformers` but not FlagEmbedding.  Same observable behaviour as FlagModel as the reference configures it
    (vector_store.py:66-77, 131-155): [CLS] pooling, L2-normalised float32 vectors, the retrieval instruction prepended
    to queries only, max_length 512, fp16 weights on a GPU.  `encode_device` / `encode_queries_device` return the vectors
    as a device tensor, so a query goes from text to the scan kernel without a host round trip.  The transformer forward
    itself is library code (HF transformers); nothing here is on the accelerated path."""

    def __init__(self, model_path: str, device="cpu", query_instruction: str = BGE_QUERY_INSTRUCTION, use_fp16: Optional[bool] = None):
        import torch
        from transformers import AutoModel, AutoTokenizer
        self.device = torch.device(device)
        self.query_instruction = query_instruction
        self.tokenizer = AutoTokenizer.from_pretrained(model_path, local_files_only=True)
        model = AutoModel.from_pretrained(model_path, local_files_only=True)
        if use_fp16 is None:
            use_fp16 = self.device.type == "cuda"
        self.model = (model.half() if use_fp16 else model.float()).to(self.device).eval()
        self.dim = int(self.model.config.hidden_size)

    def encode_device(self, texts: Sequence[str], batch_size: int = 64, max_length: int = 512, **_):
        import torch
        out = []
        with torch.no_grad():
            for i in range(0, len(texts), batch_size):
                batch = self.tokenizer(list(texts[i:i + batch_size]), padding=True, truncation=True, max_length=max_length,
                                       return_tensors="pt").to(self.device)
                cls = self.model(**batch).last_hidden_state[:, 0]
                out.append(torch.nn.functional.normalize(cls.float(), dim=-1))
        return torch.cat(out) if out else torch.zeros((0, self.dim), device=self.device)

    def encode_queries_device(self, queries: Sequence[str], **kw):
        return self.encode_device([self.query_instruction + q for q in queries], **kw)

    def encode(self, texts, **kw) -> np.ndarray:
        single = isinstance(texts, str)
        v = self.encode_device([texts] if single else list(texts), **kw).cpu().numpy()
        return v[0] if single else v

    def encode_queries(self, queries, **kw) -> np.ndarray:
        single = isinstance(queries, str)
        v = self.encode_queries_device([queries] if single else list(queries), **kw).cpu().numpy()
        return v[0] if single else v


_dense_encoder_factory: Optional[Callable] = None
_token_encoder_factory: Optional[Callable] = None


def register_dense_encoder(factory: Optional[Callable]) -> None:
    """factory(model_name, device) -> object with encode(texts) / encode_queries(texts) -> float32 [n, d]."""
    global _dense_encoder_factory
    _dense_encoder_factory = factory


def register_token_encoder(factory: Optional[Callable]) -> None:
    """factory(model_name, device) -> object with encode_query(text) -> [Lq, 128], encode_doc(text) -> [Ld, 128]."""
    global _token_encoder_factory
    _token_encoder_factory = factory


def make_dense_encoder(model_name: str, device):
    if _dense_encoder_factory is not None:
        return _dense_encoder_factory(model_name, device)
    try:
        from FlagEmbedding import FlagModel
    except ImportError as e:
        import os
        if os.path.isdir(str(model_name)):
            # a checkpoint directory on disk and no FlagEmbedding: HF transformers on the index's own device
            return TransformersBgeEncoder(str(model_name), device)
        raise RuntimeError(
            "no dense encoder: FlagEmbedding is not installed, the embedding model is not a local checkpoint directory and "
            "no encoder was registered (legal_rag_b200.retrieval.encoders.register_dense_encoder)") from e
    import torch
    return FlagModel(model_name, query_instruction_for_retrieval=BGE_QUERY_INSTRUCTION,
                     use_fp16=torch.cuda.is_available(), device=device)      # vector_store.py:70-75


def make_token_encoder(model_name: str, device):
    if _token_encoder_factory is not None:
        return _token_encoder_factory(model_name, device)
    raise RuntimeError(
        "no ColBERT token encoder registered (legal_rag_b200.retrieval.encoders.register_token_encoder); "
        "loading colbert-ai checkpoints is outside the accelerated path")
