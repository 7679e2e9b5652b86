"""Run an UNMODIFIED reference script with the B200 `Indexer` behind `src.index`.

    cd /path/to/czech-contriever            # the reference checkout (its own `src` package)
    PYTHONPATH=/path/to/czech-contriever_b200 python -m b2ip.dropin passage_retrieval.py \\
        --passages psgs.tsv --passages_embeddings "emb/passages_*" --n_docs 100 ...

The reference's scripts do `import src.index` and build `src.index.Indexer(...)`
(passage_retrieval.py:21,157).  Their own directory comes first on sys.path, so a second `src`
package on PYTHONPATH would never be found -- instead `install()` imports the reference's `src`
package and replaces its `index` submodule with one whose `Indexer` is b2ip's, before the
script runs.  faiss does not need to be installed (the reference's src/index.py is never
executed).
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types


def install(package: str = "src") -> types.ModuleType:
    """Makes `<package>.index.Indexer` the B200 Indexer.  Returns the installed module."""
    from . import indexer
    mod = types.ModuleType(f"{package}.index")
    mod.__doc__ = "b2ip drop-in for the reference's src/index.py"
    mod.Indexer = indexer.Indexer
    mod.__file__ = indexer.__file__
    sys.modules[f"{package}.index"] = mod
    try:
        pkg = importlib.import_module(package)      # the reference's package, if importable
        setattr(pkg, "index", mod)
    except ImportError:
        pass
    return mod


def install_beir() -> bool:
    """eval_beir.py / train.py never reach `Indexer`: src/beir_utils.py:14 does
    `from beir.retrieval.search.dense import DenseRetrievalExactSearch` and builds it at :167.
    When beir is importable, that name is rebound to the B200 searcher (b2ip/beir_search.py)
    before the script imports src.beir_utils.  Returns False when beir is not installed."""
    try:
        dense = importlib.import_module("beir.retrieval.search.dense")
    except ImportError:
        return False
    from .beir_search import DenseRetrievalExactSearch
    dense.DenseRetrievalExactSearch = DenseRetrievalExactSearch
    return True


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m b2ip.dropin <reference script.py> [script args...]")
    script = argv[0]
    script_dir = os.path.dirname(os.path.abspath(script))
    if script_dir not in sys.path:
        sys.path.insert(0, script_dir)              # what `python script.py` does
    install()
    install_beir()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
