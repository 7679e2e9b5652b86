"""Shared test helpers: seeded synthetic data (L2-normalised Gaussian rows,
corpus seed 1234, query seed 4321) and a restatement of the reference driver's ingest loop."""
import numpy as np


def synth(n, d, seed, normalize=True, dtype=np.float32):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(dtype)


def ingest_like_reference_driver(index, shards, indexing_batch_size):
    """What passage_retrieval.py:65-91 (index_encoded_data / add_embeddings) does with the
    (ids, embeddings) tuples unpickled from the shard files, in sorted file order: keep a
    running buffer, hand `index.index_data` full batches, then the remainder."""
    buf_ids, buf_emb = [], None
    for ids, emb in shards:
        buf_emb = emb if buf_emb is None or buf_emb.size == 0 else np.vstack((buf_emb, emb))
        buf_ids.extend(ids)
        while buf_emb.shape[0] > indexing_batch_size:
            index.index_data(buf_ids[:indexing_batch_size], buf_emb[:indexing_batch_size])
            buf_ids, buf_emb = buf_ids[indexing_batch_size:], buf_emb[indexing_batch_size:]
    while buf_emb is not None and buf_emb.shape[0] > 0:
        n = min(indexing_batch_size, buf_emb.shape[0])
        index.index_data(buf_ids[:n], buf_emb[:n])
        buf_ids, buf_emb = buf_ids[n:], buf_emb[n:]
