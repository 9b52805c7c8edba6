"""Longer version of tests/test_gpu_fuzz.py: python scripts/fuzz_gpu.py [n_cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lotus_b200 import _native as nv  # noqa: E402
from test_gpu_fuzz import one_case  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
rng = np.random.default_rng(seed)
bad = 0
for c in range(n_cases):
    try:
        one_case(nv, rng, c)
    except AssertionError as e:
        bad += 1
        print("FAIL", e, flush=True)
print(f"fuzz: {n_cases - bad}/{n_cases} cases exact", flush=True)
sys.exit(1 if bad else 0)
