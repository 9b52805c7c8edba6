"""Result carrier of the vector-store boundary. Mirrors lotus/types.py:232-235 (`RMOutput`)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any


@dataclass
class RMOutput:
    # FaissVS returns ndarray float32 [Q,K] and ndarray int64 [Q,K] (or list-of-lists when ids is given,
    # lotus/vector_store/faiss_vs.py:72,75); both are only ever indexed / iterated by the operators.
    distances: Any
    indices: Any
