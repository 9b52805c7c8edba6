"""Retrieval-model boundary. Mirrors lotus/models/rm.py:10-85: `_embed`, `__call__`, and
`convert_query_to_query_vector` which passes ndarrays through untouched (that is how precomputed embeddings
enter the path) and embeds str / list / Series."""
from __future__ import annotations

import zlib
from abc import ABC, abstractmethod
from typing import Any, Mapping

import numpy as np
import pandas as pd


class RM(ABC):
    def __init__(self) -> None:
        pass

    @abstractmethod
    def _embed(self, docs: list[str]) -> np.ndarray:
        ...

    def __call__(self, docs: list[str]) -> np.ndarray:
        return self._embed(docs)

    def convert_query_to_query_vector(self, queries: Any) -> np.ndarray:
        if isinstance(queries, str):
            queries = [queries]
        if isinstance(queries, np.ndarray):
            return queries
        try:  # device hand-off: torch tensors pass through like ndarrays
            import torch
            if isinstance(queries, torch.Tensor):
                return queries
        except Exception:  # pragma: no cover
            pass
        if isinstance(queries, pd.Series):
            queries = queries.tolist()
        return self._embed(queries)


class TableRM(RM):
    """Precomputed embeddings keyed by document text (BASELINE configs use precomputed embeddings; the text
    encoders of the reference — SentenceTransformersRM / LiteLLMRM — are out of scope)."""

    def __init__(self, table: Mapping[str, np.ndarray]):
        super().__init__()
        self.table = table

    def _embed(self, docs: list[str]) -> np.ndarray:
        return np.stack([np.asarray(self.table[d], dtype=np.float32) for d in docs]).astype(np.float32)


class HashRM(RM):
    """Deterministic pseudo-embedding of arbitrary text (seeded by crc32 of the string), L2-normalised like
    SentenceTransformersRM(normalize_embeddings=True) (lotus/models/sentence_transformers_rm.py:30,70-72).
    Test/bench plumbing only: it carries no semantics."""

    def __init__(self, dim: int = 384):
        super().__init__()
        self.dim = dim

    def _embed(self, docs: list[str]) -> np.ndarray:
        out = np.empty((len(docs), self.dim), dtype=np.float32)
        for i, d in enumerate(docs):
            rng = np.random.default_rng(zlib.crc32(str(d).encode("utf-8")))
            v = rng.standard_normal(self.dim).astype(np.float32)
            out[i] = v / np.linalg.norm(v)
        return out
