from __future__ import annotations

import warnings
from typing import Any

import pandas as pd

from ..rm import RM
from ..settings import settings
from ..vs import VS

_MSG = ("The retrieval model must be an instance of RM, and the vector store must be an instance of VS. "
        "Please configure a valid retrieval model or vector store using lotus.settings.configure()")


def active_settings():
    """The settings object the operators read: the real `lotus.settings` when lotus is importable, else ours."""
    try:
        import lotus  # type: ignore
        return lotus.settings
    except Exception:
        return settings


def resolve_rm_vs(strict: bool = False):
    s = active_settings()
    rm, vs = s.rm, s.vs
    if strict:
        ok_rm = isinstance(rm, RM) or type(rm).__name__ != "NoneType" and hasattr(rm, "convert_query_to_query_vector")
        ok_vs = isinstance(vs, VS) or (vs is not None and hasattr(vs, "load_index") and callable(vs))
        if not (ok_rm and ok_vs):
            raise ValueError(_MSG)
    elif rm is None or vs is None:
        raise ValueError(_MSG)
    return rm, vs


_REGISTRY: dict[str, type] = {}


def _register_now(name: str, cls: type):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pd.api.extensions.register_dataframe_accessor(name)(cls)


def register(name: str):
    """pd.api.extensions.register_dataframe_accessor without the override warning (the reference registers the same names)."""
    def deco(cls):
        _REGISTRY[name] = cls
        return _register_now(name, cls)
    return deco


def register_all() -> list[str]:
    """(Re-)register every accessor of this package. `import lotus` registers the reference's classes under the same names
    (lotus/sem_ops/*.py), so whoever imports lotus AFTER lotus_b200 gets the reference's N^2 sem_dedup and K-doubling
    sem_search back; `lotus_b200.install()` calls this after importing lotus."""
    for name, cls in _REGISTRY.items():
        _register_now(name, cls)
    return sorted(_REGISTRY)


def validate_df(obj: Any) -> None:
    if not isinstance(obj, pd.DataFrame):
        raise AttributeError("Must be a DataFrame")
