"""`src.index` -- the reference's module path for its retrieval index (reference src/index.py).

Copy this file over the reference's `src/index.py` (with `czech-contriever_b200/` on
PYTHONPATH), or leave the reference untouched and start its scripts through
`python -m b2ip.dropin passage_retrieval.py ...`: either way `src.index.Indexer` becomes the
B200 engine's `Indexer` (b2ip/indexer.py), same API as reference src/index.py:15-73.
"""
from b2ip.indexer import Indexer  # noqa: F401
