"""Settings singleton with the reference's semantics (lotus/settings.py:15-76): `configure(**kw)` rejects unknown
keys with ValueError, `context(**kw)` is a ContextVar overlay that nests and restores on exit.
Only the fields the embedding-similarity path reads are meaningful here (rm, vs, reranker, enable_cache, lm)."""
from __future__ import annotations

from contextlib import contextmanager
from contextvars import ContextVar
from typing import Any, Generator

_settings_context: ContextVar[dict[str, Any] | None] = ContextVar("_b200_settings_context", default=None)


class Settings:
    lm: Any | None = None
    rm: Any | None = None
    helper_lm: Any | None = None
    reranker: Any | None = None
    vs: Any | None = None
    enable_cache: bool = False
    parallel_groupby_max_threads: int = 8

    def __getattribute__(self, name: str) -> Any:
        annotations = object.__getattribute__(self, "__class__").__annotations__
        if name in annotations:
            ctx = _settings_context.get()
            if ctx is not None and name in ctx:
                return ctx[name]
        return object.__getattribute__(self, name)

    def configure(self, **kwargs: Any) -> None:
        for key, value in kwargs.items():
            if not hasattr(self, key):
                raise ValueError(f"Invalid setting: {key}")
            setattr(self, key, value)

    @contextmanager
    def context(self, **kwargs: Any) -> Generator["Settings", None, None]:
        for key in kwargs:
            if not hasattr(self, key):
                raise ValueError(f"Invalid setting: {key}")
        current = _settings_context.get() or {}
        token = _settings_context.set({**current, **kwargs})
        try:
            yield self
        finally:
            _settings_context.reset(token)

    def __str__(self) -> str:
        return str(vars(self))


settings = Settings()
