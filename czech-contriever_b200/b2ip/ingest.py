"""Streaming ingest of embedding shards.

`index_encoded_data` is a drop-in for the reference driver's function of the same name
(passage_retrieval.py:65-81): same arguments, same prints, and it hands `index.index_data`
exactly the same sequence of (ids, embeddings) batches -- consecutive `indexing_batch_size`
slices of the shard files concatenated in the given order, then the remainder.  What changes
is how the batches are produced:

  * the reference re-copies its whole pending buffer with `np.vstack` for every file
    (passage_retrieval.py:73); here a batch is a zero-copy view when it lies inside one file
    and one concatenation of at most `indexing_batch_size` rows when it spans files;
  * the next shard file is unpickled on a background thread while the current one is being
    copied to the GPU (`b2ip_add` itself double-buffers pageable host rows through pinned
    staging, csrc/b2ip_api.cu::host_to_device).
"""
from __future__ import annotations

import pickle
import queue
import threading
from typing import Callable, Iterable, Iterator, List, Sequence, Tuple, TypeVar

import numpy as np

T = TypeVar("T")
_END = object()


def prefetch(items: Iterable[T], depth: int = 1) -> Iterator[T]:
    """Iterates `items` on a background thread, `depth` items ahead of the consumer.
    Exceptions raised by the producer are re-raised in the consumer."""
    q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()

    def run():
        try:
            for it in items:
                while not stop.is_set():
                    try:
                        q.put(it, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put(_END)
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            q.put(e)

    t = threading.Thread(target=run, daemon=True)
    t.start()
    try:
        while True:
            it = q.get()
            if it is _END:
                return
            if isinstance(it, BaseException):
                raise it
            yield it
    finally:
        stop.set()


def load_shard(file_path: str) -> Tuple[list, np.ndarray]:
    """One embedding shard: pickle of (ids, embeddings [n,d]) as written by
    generate_passage_embeddings.py:94-95."""
    with open(file_path, "rb") as fin:
        ids, embeddings = pickle.load(fin)
    return ids, embeddings


def iter_batches(shards: Iterable[Tuple[Sequence, np.ndarray]], batch: int
                 ) -> Iterator[Tuple[list, np.ndarray]]:
    """Consecutive `batch`-row slices of the concatenated shards, then the remainder."""
    pend_ids: List[Sequence] = []
    pend_emb: List[np.ndarray] = []
    pending = 0

    def take(n):
        nonlocal pending
        ids_out: list = []
        emb_out: List[np.ndarray] = []
        need = n
        while need > 0:
            ids, emb = pend_ids[0], pend_emb[0]
            m = min(need, emb.shape[0])
            ids_out.extend(ids[:m])
            emb_out.append(emb[:m])
            if m == emb.shape[0]:
                pend_ids.pop(0); pend_emb.pop(0)
            else:
                pend_ids[0], pend_emb[0] = ids[m:], emb[m:]
            need -= m
        pending -= n
        return ids_out, (emb_out[0] if len(emb_out) == 1 else np.concatenate(emb_out, axis=0))

    for ids, emb in shards:
        if len(ids) != emb.shape[0]:
            raise ValueError(f"shard has {len(ids)} ids for {emb.shape[0]} rows")
        if emb.shape[0] == 0:
            continue
        pend_ids.append(ids); pend_emb.append(emb)
        pending += emb.shape[0]
        while pending > batch:                      # strict, as passage_retrieval.py:75
            yield take(batch)
    while pending > 0:
        yield take(min(batch, pending))


def index_encoded_data(index, embedding_files: Sequence[str], indexing_batch_size: int,
                       loader: Callable[[str], Tuple[list, np.ndarray]] = load_shard) -> None:
    """Drop-in for passage_retrieval.py:65-81 (see the module docstring)."""
    def shards():
        for file_path in embedding_files:
            print(f"Loading file {file_path}")
            yield loader(file_path)

    for ids, emb in iter_batches(prefetch(shards(), depth=1), indexing_batch_size):
        index.index_data(ids, emb)
    print("Data indexing completed.")
