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


def register(name: str):
    """pd.api.extensions.register_dataframe_accessor without the override warning (the reference registers the same names)."""
    def deco(cls):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return pd.api.extensions.register_dataframe_accessor(name)(cls)
    return deco


def validate_df(obj: Any) -> None:
    if not isinstance(obj, pd.DataFrame):
        raise AttributeError("Must be a DataFrame")
