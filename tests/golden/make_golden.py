"""Generates tests/golden/flatip_small.npz: seeded inputs + the fp64 brute-force answer
(score desc, row asc) for k in {1,10,100}.  The reference holds no golden vectors for this
path (no tests at all) and faiss is not installable here, so the pinned answer
is the fp64 truth; both the CPU oracle (-m "not gpu") and the CUDA path (-m gpu) are checked
against it.      python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flatip_oracle as fo  # noqa: E402

rng = np.random.default_rng(20261018)
corpus = rng.standard_normal((3000, 96)).astype(np.float32)
corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
corpus[1500:1510] = corpus[20:30]          # a few exact duplicate rows (ties)
queries = rng.standard_normal((48, 96)).astype(np.float32)
queries /= np.linalg.norm(queries, axis=1, keepdims=True)
out = {"corpus": corpus.astype(np.float16), "queries": queries.astype(np.float16)}
# stored as fp16 to keep the fixture small; the answer is computed on the fp16-rounded values
c32, q32 = out["corpus"].astype(np.float32), out["queries"].astype(np.float32)
for k in (1, 10, 100):
    D, I = fo.brute_force_f64(q32, c32, k)
    out[f"D_k{k}"] = D.astype(np.float32)
    out[f"I_k{k}"] = I
np.savez_compressed(os.path.join(os.path.dirname(__file__), "flatip_small.npz"), **out)
print("wrote flatip_small.npz")
