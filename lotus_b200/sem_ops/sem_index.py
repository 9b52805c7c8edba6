"""df.sem_index(col, index_dir) — lotus/sem_ops/sem_index.py:61-77: embed the column with settings.rm, build and
persist the index with settings.vs, record attrs["index_dirs"][col]."""
from __future__ import annotations

from typing import Any

import pandas as pd

from ..cache import operator_cache
from ._common import register, resolve_rm_vs, validate_df


@register("sem_index")
class SemIndexDataframe:
    def __init__(self, pandas_obj: Any) -> None:
        validate_df(pandas_obj)
        self._obj = pandas_obj
        # the reference resets attrs["index_dirs"] = {} here (sem_index.py:54); pandas >= 3 no longer caches
        # accessor objects, so that would wipe earlier entries on every access -> setdefault
        self._obj.attrs.setdefault("index_dirs", {})

    @operator_cache
    def __call__(self, col_name: str, index_dir: str) -> pd.DataFrame:
        rm, vs = resolve_rm_vs()
        embeddings = rm(self._obj[col_name].tolist())
        vs.index(self._obj[col_name], embeddings, index_dir)
        self._obj.attrs["index_dirs"][col_name] = index_dir
        return self._obj
