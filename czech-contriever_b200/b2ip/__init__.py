"""b2ip -- B200-native exact inner-product top-k search (host side of libb2ip.so).

Drop-in for the faiss.IndexFlatIP path of Ajchler/czech-contriever: `src.index.Indexer`
(next to this package) keeps the reference's API (src/index.py:15-73) on top of `Engine`.
"""
from ._lib import LIB_PATH, MAX_K, SYMBOLS, B2ipError, load, load_hostmap  # noqa: F401
from .engine import Engine, merge_topk  # noqa: F401
from .faiss_io import read_flat_ip_header, stream_flat_ip_rows, write_flat_ip  # noqa: F401
from .ingest import index_encoded_data, iter_batches, prefetch  # noqa: F401
from .multi import MultiGpuEngine, SegmentMap  # noqa: F401
from .sharded import ShardedIndex, shard_bounds, weighted_shard_bounds  # noqa: F401
from . import beir_search  # noqa: F401,E402  (BEIR-compatible searcher)
