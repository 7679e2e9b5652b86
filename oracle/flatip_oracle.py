"""ctypes front-end of oracle/flatip_oracle.c plus the fp64 brute force and the tie-aware
comparator the parity tests use.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py, never by the product package.

PARITY UNPINNED (see flatip_oracle.c): faiss-cpu==1.8.0 (reference environment.yml:138) is
not installable offline and the reference holds no golden vectors for this path.  The
restatement is pinned against `brute_force_f64` below instead.  `real_faiss()` is the hook
that replaces it with the real library wherever `import faiss` succeeds.

`OracleIndexer` restates the reference's `Indexer` (src/index.py:15-73) over the oracle
search so that parity tests can be written as "same calls on both objects".
"""
from __future__ import annotations

import ctypes
import glob
import os
import pickle
import struct
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libflatip_oracle.so")
_SRC = os.path.join(_HERE, "flatip_oracle.c")
_lib = None
_blas = None  # (ctypes lib, fn pointer as c_void_p, int width, description)


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/_build/ (x86-64-v3 flags so the .so runs on any
    box of the pool, not only the CPU it was built on)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        cmd = ["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-shared", "-fPIC", "-std=c11",
               "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        # The handler loops are OpenMP, the sgemm is numpy's pthread OpenBLAS: idle OpenMP
        # workers must sleep, not spin, or they starve the BLAS threads (measured 9.5 ->
        # 310 GFLOP/s on 8 cores).  libgomp reads these when it is first loaded.
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        os.environ.setdefault("GOMP_SPINCOUNT", "0")
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_flatip_search.restype = ctypes.c_int
        _lib.oracle_flatip_search.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _find_blas():
    """Locate a Fortran sgemm_ the way faiss links one: numpy's bundled OpenBLAS (ILP64
    symbols `scipy_sgemm_64_`).  Returns None when absent -> internal C loops."""
    global _blas
    if _blas is None:
        _blas = False
        pats = [os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so")]
        for p in pats:
            for so in sorted(glob.glob(p)):
                try:
                    lib = ctypes.CDLL(so)
                    fn = ctypes.cast(getattr(lib, "scipy_sgemm_64_"), ctypes.c_void_p)
                    desc = "OpenBLAS(numpy-bundled, ILP64) " + os.path.basename(so)
                    try:
                        lib.scipy_openblas_get_config64_.restype = ctypes.c_char_p
                        desc = lib.scipy_openblas_get_config64_().decode()
                    except Exception:
                        pass
                    _blas = (lib, fn, 64, desc)
                    return _blas
                except (OSError, AttributeError):
                    continue
    return _blas or None


def blas_description() -> str:
    b = _find_blas()
    return b[3] if b else "none (internal C loops)"


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def real_faiss():
    """The reference's own library when it is importable (faiss-cpu==1.8.0, reference
    environment.yml:138), else None.  It is NOT installable offline in this image (SURVEY.md
    8c); the day `import faiss` works, `search` / `make_index` below -- and with them every
    parity test, smoke() and both CPU legs of bench.py -- switch to it, and tests/test_oracle.py
    additionally checks the restatement and the index.faiss writer against it.
    B2IP_ORACLE_FAISS=0 keeps the restatement."""
    if os.environ.get("B2IP_ORACLE_FAISS", "1") == "0":
        return None
    try:
        import faiss
        return faiss
    except ImportError:
        return None


def backend_kind() -> str:
    """bench.py's `cpu_baseline.kind`: "reference" = real faiss, "port" = the restatement."""
    return "reference" if real_faiss() is not None else "port"


def backend_description() -> str:
    f = real_faiss()
    if f is not None:
        return f"faiss {getattr(f, '__version__', '?')} IndexFlatIP.search, {f.omp_get_max_threads()} OpenMP threads"
    return ("faiss-IndexFlatIP-equivalent CPU restatement (faiss-cpu 1.8.0 not installable offline), BLAS: "
            + blas_description())


def restatement_search(queries: np.ndarray, corpus: np.ndarray, k: int, use_blas: bool = True
                       ) -> Tuple[np.ndarray, np.ndarray]:
    """oracle/flatip_oracle.c: returns (D float32 [nq,k] descending, I int64 [nq,k]);
    (-FLT_MAX, -1) padding when the corpus has fewer than k rows."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    x = np.ascontiguousarray(corpus, dtype=np.float32)
    assert q.ndim == 2 and x.ndim == 2 and q.shape[1] == x.shape[1], (q.shape, x.shape)
    assert k > 0
    nq, d = q.shape
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    blas = _find_blas() if use_blas else None
    rc = _load().oracle_flatip_search(
        q.ctypes.data, x.ctypes.data, d, nq, x.shape[0], k, D.ctypes.data, I.ctypes.data,
        blas[1] if blas else None, blas[2] if blas else 0)
    if rc != 0:
        raise RuntimeError(f"oracle_flatip_search failed rc={rc}")
    return D, I


class FlatIndex:
    """`faiss.IndexFlatIP(d)` + `add(corpus)` once, `search` many times (reference
    src/index.py:21,30,42): real faiss when importable, else the restatement over the caller's
    array (no copy)."""

    def __init__(self, corpus: np.ndarray):
        self.corpus = corpus
        self._faiss_index = None
        f = real_faiss()
        if f is not None:
            self._faiss_index = f.IndexFlatIP(int(corpus.shape[1]))
            self._faiss_index.add(np.ascontiguousarray(corpus, dtype=np.float32))

    def search(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        if self._faiss_index is not None:
            return self._faiss_index.search(np.ascontiguousarray(queries, dtype=np.float32), int(k))
        return restatement_search(queries, self.corpus, k)


def make_index(corpus: np.ndarray) -> FlatIndex:
    return FlatIndex(corpus)


def search(queries: np.ndarray, corpus: np.ndarray, k: int, use_blas: bool = True
           ) -> Tuple[np.ndarray, np.ndarray]:
    """The oracle: `faiss.IndexFlatIP(d).add(corpus).search(queries, k)` -- real faiss when it
    is importable, else the restatement."""
    if use_blas and real_faiss() is not None:
        return FlatIndex(corpus).search(queries, k)
    return restatement_search(queries, corpus, k, use_blas)


def brute_force_f64(queries: np.ndarray, corpus: np.ndarray, k: int,
                    block: int = 256) -> Tuple[np.ndarray, np.ndarray]:
    """fp64 truth: full scores in float64, stable sort (score desc, row asc).
    Returns (D64 [nq,k'], I [nq,k']) with k' = min(k, N)."""
    q = np.asarray(queries, dtype=np.float64)
    x = np.asarray(corpus, dtype=np.float64)
    n = x.shape[0]
    kk = min(k, n)
    D = np.empty((q.shape[0], kk), dtype=np.float64)
    I = np.empty((q.shape[0], kk), dtype=np.int64)
    for s in range(0, q.shape[0], block):
        sc = q[s:s + block] @ x.T
        order = np.argsort(-sc, axis=1, kind="stable")[:, :kk]
        I[s:s + block] = order
        D[s:s + block] = np.take_along_axis(sc, order, axis=1)
    return D, I


def compare_topk(D_a, I_a, D_b, I_b, queries, corpus, rtol: float = 1e-5,
                 atol_scale: float = 1e-2) -> dict:
    """north_star acceptance: per query the top-k id SETS must be identical, except that ids
    present on one side only must be ties -- their scores within `rtol` relative of the
    k-th score; score vectors must agree within `rtol` relative.
    (a = implementation under test, b = oracle / truth).

    "Relative" is taken against max(|score|, atol_scale * ||q||*max||x||): an fp32 dot
    product's error is relative to sum|q_i x_i|, not to a result that cancelled to ~0, so
    scores below 1 % of the norm product (only reachable when k ~ N) get an absolute
    floor of rtol*atol_scale*||q||*max||x|| (1e-7 of the norm product) instead.
    Raises AssertionError with the first offending query; returns counters otherwise."""
    D_a = np.asarray(D_a); D_b = np.asarray(D_b); I_a = np.asarray(I_a); I_b = np.asarray(I_b)
    assert D_a.shape == D_b.shape and I_a.shape == I_b.shape, (D_a.shape, D_b.shape)
    nq, k = I_a.shape
    q64 = np.asarray(queries, dtype=np.float64)
    x64 = np.asarray(corpus, dtype=np.float64)
    floor = atol_scale * np.linalg.norm(q64, axis=1) * (
        np.sqrt((x64 * x64).sum(axis=1).max()) if x64.shape[0] else 0.0)
    n_tie_queries = 0
    scale = np.maximum(np.abs(D_b.astype(np.float64)), np.abs(D_a.astype(np.float64)))
    scale = np.maximum(scale, floor[:, None])
    valid = (I_b >= 0) & (I_a >= 0)
    assert np.array_equal(I_a >= 0, I_b >= 0), "padding (-1) pattern differs"
    if (~valid).any():
        assert np.array_equal(D_a[~valid], D_b[~valid]), "padding scores differ"
    err = np.abs(D_a.astype(np.float64) - D_b.astype(np.float64))
    bad = valid & (err > rtol * np.maximum(scale, 1e-30))
    if bad.any():
        qi, j = np.argwhere(bad)[0]
        raise AssertionError(f"score mismatch q={qi} rank={j}: {D_a[qi, j]!r} vs {D_b[qi, j]!r}")
    for qi in range(nq):
        a = set(I_a[qi][I_a[qi] >= 0].tolist()); b = set(I_b[qi][I_b[qi] >= 0].tolist())
        if a == b:
            continue
        n_tie_queries += 1
        kth = float(D_b[qi][I_b[qi] >= 0][-1])
        for r in (a ^ b):
            s = float(q64[qi] @ x64[r])
            assert abs(s - kth) <= rtol * max(abs(kth), abs(s), floor[qi], 1e-30), (
                f"q={qi}: id {r} on one side only with score {s!r}, k-th score {kth!r} -- not a tie")
    return {"queries": nq, "queries_with_tie_swaps": n_tie_queries}


# ------------------------------------------------------------------ index.faiss format
# faiss 1.8.0 impl/index_write.cpp for IndexFlatIP (restated from the published format;
# no faiss here to cross-check -> SURVEY.md 8b): fourcc "IxFI", int32 d, int64 ntotal,
# int64 dummy(1<<20) x2, uint8 is_trained, int32 metric_type(0 = inner product),
# uint64 count = ntotal*d, then float32[count] row-major, all little-endian.
_HDR = struct.Struct("<4siqqqBi")


def write_index_flat_ip(path: str, rows: np.ndarray) -> None:
    rows = np.ascontiguousarray(rows, dtype="<f4")
    n, d = rows.shape
    with open(path, "wb") as f:
        f.write(_HDR.pack(b"IxFI", d, n, 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", n * d))
        f.write(rows.tobytes())


def read_index_flat_ip(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        cc, d, n, _, _, trained, metric = _HDR.unpack(f.read(_HDR.size))
        if cc != b"IxFI" or metric != 0:
            raise ValueError(f"not an IndexFlatIP file: fourcc={cc!r} metric={metric}")
        (count,) = struct.unpack("<Q", f.read(8))
        if count != n * d:
            raise ValueError("corrupt index.faiss: count != ntotal*d")
        return np.frombuffer(f.read(4 * count), dtype="<f4").reshape(n, d).copy()


class OracleIndexer:
    """Restatement of reference `Indexer` (src/index.py:15-73) on the CPU oracle."""

    def __init__(self, vector_sz, n_subquantizers=0, n_bits=8):
        if n_subquantizers > 0:
            raise NotImplementedError("IndexPQ branch (src/index.py:18-19) is out of scope")
        self.d = vector_sz
        self.rows = np.empty((0, vector_sz), dtype=np.float32)
        self.index_id_to_db_id = []

    def index_data(self, ids, embeddings):  # src/index.py:25-32
        self.index_id_to_db_id.extend(ids)
        self.rows = np.concatenate([self.rows, embeddings.astype("float32")], axis=0)

    def search_knn(self, query_vectors, top_docs, index_batch_size=2048
                   ) -> List[Tuple[List[object], np.ndarray]]:  # src/index.py:34-46
        query_vectors = query_vectors.astype("float32")
        result = []
        nbatch = (len(query_vectors) - 1) // index_batch_size + 1
        for b in range(nbatch):
            q = query_vectors[b * index_batch_size:(b + 1) * index_batch_size]
            scores, indexes = search(q, self.rows, top_docs)
            db_ids = [[str(self.index_id_to_db_id[i]) for i in row] for row in indexes]
            result.extend([(db_ids[i], scores[i]) for i in range(len(db_ids))])
        return result

    def serialize(self, dir_path):  # src/index.py:48-55
        write_index_flat_ip(os.path.join(dir_path, "index.faiss"), self.rows)
        with open(os.path.join(dir_path, "index_meta.faiss"), "wb") as f:
            pickle.dump(self.index_id_to_db_id, f)

    def deserialize_from(self, dir_path):  # src/index.py:57-68
        self.rows = read_index_flat_ip(os.path.join(dir_path, "index.faiss"))
        with open(os.path.join(dir_path, "index_meta.faiss"), "rb") as f:
            self.index_id_to_db_id = pickle.load(f)
        assert len(self.index_id_to_db_id) == self.rows.shape[0]
