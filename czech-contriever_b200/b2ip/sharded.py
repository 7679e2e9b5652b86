"""Row-sharded search across the GPUs of one box: one process per GPU (`torchrun`), each rank
owns a set of corpus segments in its own `Engine`, queries are replicated, every rank computes
a local top-k and the one exchange step is an all-gather of the [nq,k] (score, global row)
blocks over NCCL/NVLink followed by the merge kernel.

With `peer_exchange` (default when torch symmetric memory can map the ranks' buffers into each
other, i.e. on one NVLink/NVSwitch box) the exchange is fused behind the search instead: the
engine's finalize kernel stores every rank's block straight into all peers' gather buffers over
NVLink, a flag is published, and each rank's merge kernel starts as soon as all flags are in
(`b2ip_search_exchange`, include/b2ip.h) -- no NCCL call and no host round trip between the
local search and the merge, which is what the latency regime (BASELINE config 5) is made of.

The reference runs this path in a single process on host cores (faiss OpenMP,
src/index.py:42); sharding is how the same search is spread over 1/2/4/8 B200s.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row shard of `rank`: [rank*ceil(N/G), min((rank+1)*ceil(N/G), N))."""
    per = -(-n_total // world) if world > 0 else n_total
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def weighted_shard_bounds(n_total: int, weights, rank: int, align: int = 256) -> Tuple[int, int]:
    """Contiguous row shard of `rank` when the ranks are not equally fast: shard sizes proportional
    to `weights` (e.g. each GPU's measured scoring rate -- under the power cap the B200s of one
    box sustain clocks a few percent apart, and a row-sharded search ends with its slowest
    rank), boundaries rounded to `align` rows.  Equal weights give `shard_bounds` up to rounding."""
    w = [max(float(x), 0.0) for x in weights]
    tot = sum(w)
    if tot <= 0 or len(w) == 0:
        return shard_bounds(n_total, max(len(w), 1), rank)
    edges, acc = [0], 0.0
    for x in w[:-1]:
        acc += x
        e = int(round(n_total * acc / tot / align)) * align
        edges.append(min(max(e, edges[-1]), n_total))
    edges.append(n_total)
    return edges[rank], edges[rank + 1]


class ShardedIndex:
    FLAG_BYTES = 256          # per parity: uint32[4 * world] flags (XF_WORDS per rank), padded

    def __init__(self, d: int, device: Optional[int] = None, engine=None, group=None,
                 merge_fn: Optional[Callable] = None, store: str = "f32",
                 peer_exchange: Optional[bool] = None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if engine is None:
            from .engine import Engine, merge_topk
            if device is None:
                device = torch.cuda.current_device()
            engine = Engine(d, device, store=store)
            merge_fn = merge_fn or merge_topk
        self.engine = engine
        self.merge_fn = merge_fn
        self.d = d
        self.segments: List[Tuple[int, int, int]] = []   # (local_start, global_start, n)
        self._n_local = 0
        self._n_seen = 0          # add_replicated: rows seen by every rank so far
        self._chunk_no = 0
        self._seg_cache = None
        if peer_exchange is None:
            peer_exchange = os.environ.get("B2IP_PEER_EXCHANGE", "1") != "0"
        self.peer_exchange = bool(peer_exchange) and self.world > 1 and self.world <= 8 \
            and self._engine_writes_in_place()
        # global threshold round of the exchange (include/b2ip.h): each rank rescores ~k/world rows
        self.global_threshold = os.environ.get("B2IP_GLOBAL_THRESHOLD", "1") != "0"
        self._x_buf = None        # symmetric buffer [gather slots | threshold rows | flag arrays] x 2 parities
        self._x_slot = 0
        self._x_thr_cap = 0
        self._x_ex = None
        self._x_seq = 0
        self.exchange_searches = 0

    # -- ingest ----------------------------------------------------------------------
    def add_local(self, rows, global_start: int) -> None:
        """Rows this rank owns; they take global ids global_start .. global_start+n-1.
        Segments must be added in increasing global order (keeps the lower-row tie rule)."""
        n = int(rows.shape[0])
        if n == 0:
            return
        if self.segments and global_start < self.segments[-1][1] + self.segments[-1][2]:
            raise ValueError("segments must be appended in increasing global row order")
        self.engine.add(rows)
        if self.segments and self.segments[-1][1] + self.segments[-1][2] == global_start:
            ls, gs, m = self.segments[-1]
            self.segments[-1] = (ls, gs, m + n)
        else:
            self.segments.append((self._n_local, int(global_start), n))
        self._n_local += n
        self._seg_cache = None
        # the engine maps local rows to global ids itself (offset or segment table): no extra
        # kernels per search, and its results can go straight to the peers
        if hasattr(self.engine, "set_row_segments"):
            self.engine.set_row_segments(self.segments)
        elif hasattr(self.engine, "set_row_offset"):
            self.engine.set_row_offset(self.segments[0][1] - self.segments[0][0]
                                       if len(self.segments) == 1 else 0)

    def add_replicated(self, rows) -> None:
        """SPMD ingest: every rank is handed the SAME chunk (the reference's `index_data`
        stream, passage_retrieval.py:65-91); chunk i is kept by rank i % world."""
        n = int(rows.shape[0])
        if self._chunk_no % self.world == self.rank:
            self.add_local(rows, self._n_seen)
        self._n_seen += n
        self._chunk_no += 1

    @property
    def ntotal_local(self) -> int:
        return self._n_local

    # -- search ----------------------------------------------------------------------
    def _to_global(self, rows_local):
        torch = self._torch
        if not self.segments or hasattr(self.engine, "set_row_segments"):
            return rows_local                          # already global
        if len(self.segments) == 1:
            if hasattr(self.engine, "set_row_offset"):
                return rows_local                      # offset already applied by the engine
            ls, gs, _ = self.segments[0]
            return torch.where(rows_local >= 0, rows_local + (gs - ls), rows_local)
        if self._seg_cache is None or self._seg_cache[0].device != rows_local.device:
            ls = torch.tensor([s[0] for s in self.segments], dtype=torch.int64, device=rows_local.device)
            delta = torch.tensor([s[1] - s[0] for s in self.segments], dtype=torch.int64,
                                 device=rows_local.device)
            self._seg_cache = (ls, delta)
        ls, delta = self._seg_cache
        seg = torch.bucketize(rows_local.clamp(min=0), ls, right=True) - 1
        return torch.where(rows_local >= 0, rows_local + delta[seg], rows_local)

    def search_local(self, queries, k: int, mode: str = "auto"):
        D, I = self.engine.search(queries, k, mode=mode)
        return D, self._to_global(I)

    def search(self, queries, k: int, mode: str = "auto"):
        """queries: [nq,d] tensor on this rank's device (identical on all ranks).
        Returns the global (scores [nq,k], rows [nq,k]) on every rank."""
        if self.world == 1:
            return self.search_local(queries, k, mode)
        D, I, _ = self._search_exchange(queries, k, mode, owner=False)
        return D, I

    def search_owned(self, queries, k: int, mode: str = "auto"):
        """The same search with the result PARTITIONED over the ranks: every rank gets the global
        top-k of the contiguous slice of queries it owns, `shard_bounds(nq, world, rank)`.
        Returns (scores [q_hi-q_lo,k], rows [q_hi-q_lo,k], (q_lo, q_hi)).  This is the scalable
        form of the exchange: a rank's block for a query crosses NVLink once (to the owner), each
        query is merged once, and a consumer on the host reads 1/world of the result from every
        GPU's copy engine in parallel."""
        nq = int(queries.shape[0])
        lo, hi = shard_bounds(nq, self.world, self.rank)
        if self.world == 1:
            D, I = self.search_local(queries, k, mode)
            return D, I, (lo, hi)
        return self._search_exchange(queries, k, mode, owner=True)

    def _search_exchange(self, queries, k: int, mode: str, owner: bool):
        nq = int(queries.shape[0])
        lo, hi = shard_bounds(nq, self.world, self.rank)
        if self.peer_exchange and mode != "exact" and nq > 0:
            try:
                ex = self._exchange_buffers(nq, int(k))
            except Exception as e:  # noqa: BLE001 - no symmetric memory on this box (same on every rank)
                import warnings
                warnings.warn(f"peer-direct exchange unavailable ({e!r}); using the NCCL all-gather path")
                self.peer_exchange = False
                ex = None
            if ex is not None:
                self._x_seq += 1
                D, I, status = self.engine.search_exchange(
                    queries, k, ex[1 if owner else 0][self._x_seq & 1], self._x_seq)
                if status == 0:
                    self.exchange_searches += 1
                    return D, I, (lo, hi)
                # some rank overflowed a candidate list (same status everywhere): all ranks repeat
                # the search below, where the exact fallback runs before the exchange
        D, I = self._search_allgather(queries, k, mode)
        if owner:
            return D[lo:hi].contiguous(), I[lo:hi].contiguous(), (lo, hi)
        return D, I, (lo, hi)

    def search_host(self, queries, k: int, out=None, mode: str = "auto"):
        """Host buffers in and out (SPMD: every rank passes the same numpy queries [nq,d],
        float16 or float32).  The queries go up through the engine's pinned staging, the global
        top-k is computed as in `search`, and this rank downloads the contiguous slice of queries
        it owns -- `shard_bounds(nq, world, rank)` -- into `out` = (D [nq,k] float32, I [nq,k]
        int64).  With `out` backed by memory the ranks share (e.g. np.memmap of one /dev/shm
        file) the union of the slices is the whole answer and the D2H runs on every GPU's copy
        engine at once.  Returns (D, I, (q_lo, q_hi))."""
        import numpy as np
        torch = self._torch
        q = np.asarray(queries)
        if q.dtype not in (np.float16, np.float32):
            q = q.astype(np.float32)
        q = np.ascontiguousarray(q)
        nq, k = int(q.shape[0]), int(k)
        dev = torch.device("cuda", self.engine.device)
        if out is None:
            out = (np.empty((nq, k), dtype=np.float32), np.empty((nq, k), dtype=np.int64))
        lo, hi = shard_bounds(nq, self.world, self.rank)
        if nq == 0:
            return out[0], out[1], (lo, hi)
        tq = torch.empty((nq, self.d), dtype=torch.float16 if q.dtype == np.float16 else torch.float32, device=dev)
        if self.world > 1 and nq >= 64 * self.world:
            # every rank uploads the slice of queries it owns, NVLink all-gathers them
            per = -(-nq // self.world)
            pad = torch.empty((self.world * per, self.d), dtype=tq.dtype, device=dev)
            if hi > lo:
                self.engine.upload(pad[lo:hi], q[lo:hi])
            self._dist.all_gather_into_tensor(pad, pad[self.rank * per:(self.rank + 1) * per], group=self.group)
            tq = pad[:nq]
        else:
            self.engine.upload(tq, q)
        D, I, _ = self.search_owned(tq, k, mode)
        if hi > lo:
            self.engine.download(D, out[0][lo:hi])
            self.engine.download(I, out[1][lo:hi])
        return out[0], out[1], (lo, hi)

    # -- peer-direct exchange --------------------------------------------------------
    def _exchange_buffers(self, nq: int, k: int):
        """Symmetric (peer-mapped) memory of the exchange, two parities, grown collectively:
        [2 x world gather slots | 2 x world x thr_cap floats (global threshold) | 2 flag arrays].
        Returns ex[owner][parity] (_lib.Exchange; owner = 0: every rank gets everything, 1: owner mode)."""
        need = (nq * k * 12 + 15) // 16 * 16
        if self._x_ex is not None and need <= self._x_slot and nq <= self._x_thr_cap:
            return self._x_ex
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import GATHER_ALL, GATHER_OWNER, Exchange
        torch, dist = self._torch, self._dist
        dev = torch.device("cuda", self.engine.device)
        # 25 % headroom (fewer collective re-allocations as nq*k creeps up), 64 KiB granules
        slot = max(1 << 16, (need + need // 4 + 65535) // 65536 * 65536, self._x_slot)
        thr_cap = max((nq + nq // 4 + 1023) // 1024 * 1024, self._x_thr_cap)
        thr_bytes = self.world * thr_cap * 4
        total = 2 * self.world * slot + 2 * thr_bytes + 2 * self.FLAG_BYTES
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)                 # nobody still uses the old buffers
        buf = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        buf.zero_()
        torch.cuda.synchronize(dev)
        hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        hdl.barrier()                                  # every rank's flags are zero before any write
        ptrs = list(hdl.buffer_ptrs)
        exs = []
        for gather_mode in (GATHER_ALL, GATHER_OWNER):
            pair = []
            for parity in (0, 1):
                ex = Exchange()
                ex.world, ex.rank, ex.slot_bytes = self.world, self.rank, slot
                ex.thr_stride, ex.gather_mode = thr_cap, gather_mode
                for p in range(self.world):
                    ex.gather[p] = ptrs[p] + parity * self.world * slot
                    if self.global_threshold:
                        ex.gthr[p] = ptrs[p] + 2 * self.world * slot + parity * thr_bytes
                    ex.flags[p] = ptrs[p] + 2 * self.world * slot + 2 * thr_bytes + parity * self.FLAG_BYTES
                pair.append(ex)
            exs.append(pair)
        self._x_buf, self._x_hdl, self._x_slot, self._x_thr_cap, self._x_ex, self._x_seq = buf, hdl, slot, thr_cap, exs, 0
        return exs

    # -- all-gather exchange -----------------------------------------------------------
    def _search_allgather(self, queries, k: int, mode: str = "auto"):
        """One exchange step: every rank owns a slot [rows int64 | scores fp32] of ONE packed
        buffer, the engine writes its local top-k straight into its slot, a single in-place
        all-gather fills the others, and the merge kernel reads the slots where they are."""
        torch, dist = self._torch, self._dist
        nq = int(queries.shape[0])
        nb = nq * k
        slot = (nb * 12 + 15) // 16 * 16
        buf = torch.empty((self.world, slot), dtype=torch.uint8, device=queries.device)
        mine = buf[self.rank]
        I_mine = mine[:nb * 8].view(torch.int64).view(nq, k)
        D_mine = mine[nb * 8:nb * 12].view(torch.float32).view(nq, k)
        if self._engine_writes_in_place():
            self.engine.search(queries, k, mode=mode, out=(D_mine, I_mine))
            Ig = self._to_global(I_mine)
            if Ig is not I_mine:
                I_mine.copy_(Ig)
        else:
            D, I = self.search_local(queries, k, mode)
            D_mine.copy_(D)
            I_mine.copy_(I)
        try:
            dist.all_gather_into_tensor(buf.view(-1), mine, group=self.group)
        except (RuntimeError, NotImplementedError):
            dist.all_gather([buf[r] for r in range(self.world)], mine.clone(), group=self.group)
        gI = buf[:, :nb * 8].view(torch.int64).view(self.world, nq, k)
        gD = buf[:, nb * 8:nb * 12].view(torch.float32).view(self.world, nq, k)
        return self.merge_fn(gD, gI, k)

    def _engine_writes_in_place(self) -> bool:
        from .engine import Engine
        return isinstance(self.engine, Engine)
