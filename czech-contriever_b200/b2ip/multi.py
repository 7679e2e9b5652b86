"""`MultiGpuEngine`: the row-sharded search of `ShardedIndex`, driven from ONE process.

The reference runs its retrieval in a single process (passage_retrieval.py never initialises
torch.distributed), so a drop-in `Indexer` cannot assume `torchrun`.  This class
gives that single process every GPU of the box: one `Engine` (= one C-ABI handle, one row
shard) per device, host threads that run the per-device searches concurrently (the ctypes call
releases the GIL), peer-to-peer copies of the per-shard [nq,k] results to the first device
and the same merge kernel (`b2ip_merge_topk_strided`) the NCCL path uses.  Same surface as
`Engine` (add / search / ntotal / export_rows / reserve / stats / close), same results: the
merge keeps score-descending order and the lower GLOBAL row among equal scores.

Every `add` splits its rows contiguously over the devices, so shards stay balanced however
the caller chunks its data; a per-engine segment table maps local rows back to global ids.
"""
from __future__ import annotations

import bisect
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .engine import Engine, merge_topk


class SegmentMap:
    """Local row -> global row for one shard: segments (local_start, global_start, n) appended
    in increasing global order."""

    def __init__(self):
        self.segments: List[Tuple[int, int, int]] = []
        self.n_local = 0
        self._cache = None

    def append(self, global_start: int, n: int) -> None:
        if n <= 0:
            return
        if self.segments and global_start < self.segments[-1][1] + self.segments[-1][2]:
            raise ValueError("segments must be appended in increasing global row order")
        if self.segments and self.segments[-1][1] + self.segments[-1][2] == global_start:
            ls, gs, m = self.segments[-1]
            self.segments[-1] = (ls, gs, m + n)
        else:
            self.segments.append((self.n_local, int(global_start), int(n)))
        self.n_local += n
        self._cache = None

    def single_offset(self) -> Optional[int]:
        """global - local when the shard is one contiguous segment (the engine can add it)."""
        if len(self.segments) == 1:
            return self.segments[0][1] - self.segments[0][0]
        return None

    def to_global(self, rows_local):
        """torch int64 tensor of local rows (-1 = padding) -> global rows."""
        import torch
        if not self.segments:
            return rows_local
        if self._cache is None or self._cache[0].device != rows_local.device:
            ls = torch.tensor([s[0] for s in self.segments], dtype=torch.int64, device=rows_local.device)
            delta = torch.tensor([s[1] - s[0] for s in self.segments], dtype=torch.int64,
                                 device=rows_local.device)
            self._cache = (ls, delta)
        ls, delta = self._cache
        seg = torch.bucketize(rows_local.clamp(min=0), ls, right=True) - 1
        return torch.where(rows_local >= 0, rows_local + delta[seg], rows_local)

    def overlaps(self, g0: int, g1: int):
        """(local_start, global_start, n) pieces of this shard inside global rows [g0, g1)."""
        starts = [s[1] for s in self.segments]
        i = max(0, bisect.bisect_right(starts, g0) - 1)
        for ls, gs, n in self.segments[i:]:
            if gs >= g1:
                break
            a, b = max(gs, g0), min(gs + n, g1)
            if a < b:
                yield ls + (a - gs), a, b - a


class MultiGpuEngine:
    def __init__(self, d: int, devices: Optional[Sequence[int]] = None, store: str = "f32",
                 shadow: Optional[str] = None):
        import torch
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        devices = [int(x) for x in devices]
        if not devices:
            raise RuntimeError("MultiGpuEngine needs at least one CUDA device (no CPU fallback)")
        self._torch = torch
        self.d, self.store, self.devices = int(d), store, devices
        self.engines = [Engine(d, dev, store=store, shadow=shadow) for dev in devices]
        self.maps = [SegmentMap() for _ in devices]
        self._n = 0
        self._pool = ThreadPoolExecutor(max_workers=len(devices))
        self._last_stats: List[dict] = []

    # -- lifetime / configuration ----------------------------------------------------
    def close(self) -> None:
        for e in getattr(self, "engines", []):
            e.close()
        if getattr(self, "_pool", None) is not None:
            self._pool.shutdown(wait=False)
            self._pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device(self) -> int:
        return self.devices[0]

    @property
    def ntotal(self) -> int:
        return self._n

    def reserve(self, n_rows: int) -> None:
        per = -(-int(n_rows) // len(self.engines))
        for e in self.engines:
            e.reserve(per + len(self.engines))

    def set_option(self, name: str, value: int) -> None:
        for e in self.engines:
            e.set_option(name, value)

    # -- data ------------------------------------------------------------------------
    def add(self, rows) -> None:
        """Rows get global ids ntotal .. ntotal+n-1; slice g of the chunk goes to device g."""
        n = int(rows.shape[0])
        G = len(self.engines)
        per = -(-n // G) if n else 0
        jobs = []
        for g, e in enumerate(self.engines):
            lo, hi = min(g * per, n), min((g + 1) * per, n)
            if hi > lo:
                self.maps[g].append(self._n + lo, hi - lo)
                jobs.append(self._pool.submit(self._add_one, g, rows[lo:hi]))
        for j in jobs:
            j.result()
        for g, e in enumerate(self.engines):
            e.set_row_segments(self.maps[g].segments)      # the engine reports global ids itself
        self._n += n

    def _add_one(self, g: int, part) -> None:
        torch = self._torch
        if type(part).__module__.startswith("torch") and part.is_cuda and part.device.index != self.devices[g]:
            part = part.to(torch.device("cuda", self.devices[g]))
        with torch.cuda.device(self.devices[g]):
            self.engines[g].add(part)

    def export_rows(self, row0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.d), dtype=np.float32)
        for g, e in enumerate(self.engines):
            for ls, gs, m in self.maps[g].overlaps(row0, row0 + n):
                out[gs - row0:gs - row0 + m] = e.export_rows(ls, m)
        return out

    # -- search ----------------------------------------------------------------------
    def _search_one(self, g: int, q_dev0, k: int, mode: str):
        torch = self._torch
        dev = torch.device("cuda", self.devices[g])
        with torch.cuda.device(dev):
            q = q_dev0 if q_dev0.device == dev else q_dev0.to(dev)      # fan-out over NVLink
            D, I = self.engines[g].search(q, k, mode=mode)
            st = self.engines[g].stats()
            torch.cuda.current_stream(dev).synchronize()
        return D, I, st

    def search(self, queries, k: int, mode: str = "auto", out=None):
        """numpy in -> numpy out, torch CUDA tensor in -> tensors on the first device."""
        torch = self._torch
        k = int(k)
        is_t = type(queries).__module__.startswith("torch")
        if is_t:
            q_dev0, q_host = queries.float().contiguous(), None
            torch.cuda.current_stream(q_dev0.device).synchronize()
            nq = q_dev0.shape[0]
        else:
            q_host = np.ascontiguousarray(np.asarray(queries), dtype=np.float32)
            if q_host.ndim != 2 or q_host.shape[1] != self.d:
                raise ValueError(f"expected [nq,{self.d}] queries, got {q_host.shape}")
            nq = q_host.shape[0]
            with torch.cuda.device(self.devices[0]):
                # one staged upload (pinned double buffering inside the library), then P2P fan-out
                q_dev0 = torch.empty((nq, self.d), dtype=torch.float32,
                                     device=torch.device("cuda", self.devices[0]))
                if nq:
                    self.engines[0].upload(q_dev0, q_host)
        dev0 = torch.device("cuda", self.devices[0])
        if nq == 0:
            D = torch.empty((0, k), dtype=torch.float32, device=dev0)
            I = torch.empty((0, k), dtype=torch.int64, device=dev0)
        else:
            parts = [self._pool.submit(self._search_one, g, q_dev0, k, mode)
                     for g in range(len(self.engines))]
            res = [p.result() for p in parts]
            self._last_stats = [r[2] for r in res]
            with torch.cuda.device(dev0):
                if len(res) == 1:
                    D, I = res[0][0], res[0][1]
                else:
                    gD = torch.stack([r[0].to(dev0) for r in res])
                    gI = torch.stack([r[1].to(dev0) for r in res])
                    D, I = merge_topk(gD, gI, k)
        if is_t:
            return D, I
        Dn, In = out if out is not None else (np.empty((nq, k), np.float32), np.empty((nq, k), np.int64))
        if nq:
            self.engines[0].download(D.contiguous(), Dn)
            self.engines[0].download(I.contiguous(), In)
        return Dn, In

    def stats(self) -> dict:
        """Per-device stats of the last search, plus sums of the additive counters."""
        if not self._last_stats:
            return {}
        agg = dict(self._last_stats[0])
        for key in ("coarse_launches", "total_launches", "coarse_flops", "candidates", "rescored",
                    "fallback_queries"):
            agg[key] = sum(s[key] for s in self._last_stats)
        for key in ("coarse_ms", "total_ms", "refresh_ms", "finalize_ms"):
            agg[key] = max(s[key] for s in self._last_stats)
        agg["ntotal"] = self._n
        agg["per_device"] = self._last_stats
        return agg
