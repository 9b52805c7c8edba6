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


class BagOfWordsRM(RM):
    """Lexical stand-in for a sentence encoder (examples only): every lower-cased word owns a fixed pseudo-random direction
    (seeded by its crc32); a text embeds as the L2-normalised sum of its words' directions, so texts sharing words are close."""

    def __init__(self, dim: int = 256):
        super().__init__()
        self.dim = dim
        self._words: dict[str, np.ndarray] = {}

    def _word(self, w: str) -> np.ndarray:
        v = self._words.get(w)
        if v is None:
            v = np.random.default_rng(zlib.crc32(w.encode("utf-8"))).standard_normal(self.dim).astype(np.float32)
            self._words[w] = v
        return v

    def _embed(self, docs: list[str]) -> np.ndarray:
        out = np.zeros((len(docs), self.dim), dtype=np.float32)
        for i, d in enumerate(docs):
            words = [w for w in "".join(c.lower() if c.isalnum() else " " for c in str(d)).split() if w]
            for w in words:
                out[i] += self._word(w)
            nrm = float(np.linalg.norm(out[i]))
            out[i] = out[i] / nrm if nrm > 0 else self._word("")
        return out

