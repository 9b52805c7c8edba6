"""`operator_cache` with the reference's pass-through behaviour (lotus/cache.py:33-100): a no-op unless
`settings.enable_cache` is set AND `settings.lm.cache` exists. The LM caches themselves are out of scope, so
with caching enabled this decorator dereferences `settings.lm.cache` exactly like the reference (and raises
AttributeError when no LM is configured, lotus/cache.py:38-41)."""
from __future__ import annotations

import hashlib
import json
from functools import wraps
from typing import Any, Callable

import pandas as pd



def operator_cache(func: Callable) -> Callable:
    @wraps(func)
    def wrapper(self, *args, **kwargs):
        from .sem_ops._common import active_settings  # the real lotus.settings when lotus is importable
        cfg = active_settings()
        model = cfg.lm
        if cfg.enable_cache and model.cache is not None:

            def serialize(value: Any) -> Any:
                if value is None or isinstance(value, (str, int, float, bool)):
                    return value
                if isinstance(value, pd.DataFrame):
                    return value.to_json(orient="split")
                if isinstance(value, (list, tuple)):
                    return [serialize(v) for v in value]
                if isinstance(value, dict):
                    return {k: serialize(v) for k, v in value.items()}
                return str(value)

            key = hashlib.sha256(json.dumps({"self": serialize(self._obj), "args": [serialize(a) for a in args],
                                             "kwargs": {k: serialize(v) for k, v in kwargs.items()}},
                                            sort_keys=True).encode()).hexdigest()
            hit = model.cache.get(key)
            if hit is not None:
                return hit
            result = func(self, *args, **kwargs)
            model.cache.insert(key, result)
            return result
        return func(self, *args, **kwargs)

    return wrapper
