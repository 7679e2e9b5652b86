"""`index.faiss` reader / writer for the one index type on the path, `IndexFlatIP`
(reference: faiss.write_index at src/index.py:53, faiss.read_index at src/index.py:62).

Layout (faiss 1.8.0 impl/index_write.cpp + io macros, little-endian, restated from the
published format -- no faiss is installable here to cross-check):
    fourcc "IxFI" | int32 d | int64 ntotal | int64 dummy=1<<20 | int64 dummy=1<<20 |
    uint8 is_trained | int32 metric_type (0 = METRIC_INNER_PRODUCT) |
    uint64 count = ntotal*d | float32[count] row-major
Rows are streamed in chunks so a 64 GB index never needs a second host copy, and the file
I/O runs one chunk ahead of / behind the GPU copies on a background thread.
"""
from __future__ import annotations

import struct
from typing import Callable, Iterator, Tuple

import numpy as np

_HDR = struct.Struct("<4siqqqBi")
_CHUNK_ROWS = 1 << 18


def write_flat_ip(path: str, d: int, ntotal: int,
                  export_rows: Callable[[int, int], np.ndarray]) -> None:
    """export_rows(row0, n) -> float32 [n,d] (e.g. Engine.export_rows)."""
    from .ingest import prefetch

    def chunks():
        for r0 in range(0, ntotal, _CHUNK_ROWS):
            yield np.ascontiguousarray(export_rows(r0, min(_CHUNK_ROWS, ntotal - r0)), dtype="<f4")

    with open(path, "wb") as f:
        f.write(_HDR.pack(b"IxFI", d, ntotal, 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", ntotal * d))
        for block in prefetch(chunks(), depth=1):      # D2H of chunk i+1 overlaps the write of chunk i
            f.write(memoryview(block).cast("B"))


def read_flat_ip_header(f) -> Tuple[int, int]:
    raw = f.read(_HDR.size)
    if len(raw) != _HDR.size:
        raise ValueError("index.faiss: truncated header")
    cc, d, ntotal, _, _, _trained, metric = _HDR.unpack(raw)
    if cc != b"IxFI":
        raise NotImplementedError(
            f"index.faiss holds index type {cc!r}; only IndexFlatIP ('IxFI') is on the B200 path")
    if metric != 0:
        raise ValueError(f"index.faiss: metric_type={metric}, expected 0 (inner product)")
    (count,) = struct.unpack("<Q", f.read(8))
    if count != ntotal * d:
        raise ValueError("index.faiss: vector count does not match ntotal*d")
    return d, ntotal


def stream_flat_ip_rows(path: str) -> Tuple[int, int, Iterator[np.ndarray]]:
    """(d, ntotal, iterator over float32 [<=chunk, d] blocks)."""
    f = open(path, "rb")
    d, ntotal = read_flat_ip_header(f)

    def blocks():
        try:
            for r0 in range(0, ntotal, _CHUNK_ROWS):
                n = min(_CHUNK_ROWS, ntotal - r0)
                buf = f.read(4 * n * d)
                if len(buf) != 4 * n * d:
                    raise ValueError("index.faiss: truncated vector data")
                yield np.frombuffer(buf, dtype="<f4").reshape(n, d)
        finally:
            f.close()

    from .ingest import prefetch
    return d, ntotal, prefetch(blocks(), depth=2)     # the read of chunk i+1 overlaps the H2D of chunk i
