"""Return types of the hot path.  When the reference package is importable its own pydantic models are
used (so hits flow into RagPipeline unchanged); otherwise field-for-field mirrors of
legalrag/schemas.py:9-32 are defined here."""
from __future__ import annotations

from typing import Any, Dict, List, Literal, Optional

try:  # drop-in inside the reference application
    from legalrag.schemas import LawChunk, RetrievalHit  # type: ignore  # noqa: F401
    USING_REFERENCE_SCHEMAS = True
except Exception:  # standalone
    from pydantic import BaseModel, ConfigDict

    USING_REFERENCE_SCHEMAS = False

    class LawChunk(BaseModel):
        id: str
        law_name: str
        chapter: Optional[str] = None
        section: Optional[str] = None
        article_no: str
        article_id: str
        text: str
        lang: Optional[str] = "zh"
        source: Optional[str] = None
        start_char: Optional[int] = None
        end_char: Optional[int] = None

    class RetrievalHit(BaseModel):
        model_config = ConfigDict(arbitrary_types_allowed=True)
        chunk: LawChunk
        score: float
        rank: Optional[int] = None
        source: Literal["retriever", "graph", "rerank"] = "retriever"
        semantic_score: Optional[float] = None
        graph_depth: Optional[int] = None
        relations: Optional[List[str]] = None
        seed_article_id: Optional[str] = None
        score_breakdown: Optional[Dict[str, Any]] = None


def chunk_from_obj(obj) -> "LawChunk":
    """dict / LawChunk / foreign pydantic model -> LawChunk (bm25_retriever.py:51-61)."""
    if isinstance(obj, LawChunk):
        return obj
    if isinstance(obj, dict):
        return LawChunk(**obj)
    try:
        return LawChunk(**obj.model_dump())
    except Exception as e:  # same failure mode as the reference
        raise RuntimeError(f"unsupported chunk format in index: {type(obj)}") from e
