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

    B2IP_DEVICE_QUERIES=1 python -m b2ip.dropin passage_retrieval.py ...

additionally replaces the script's own `embed_queries` by `embed_queries_device`, which leaves the
encoder's output on the GPU; `Indexer.search_knn` takes that CUDA tensor as it is (SURVEY 8f N4:
no device -> host -> device round trip of the queries).  B2IP_DEVICES=all row-shards the index
over every GPU of the box from this one process.
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


def embed_queries_device(args, queries, model, tokenizer):
    """Device-resident replacement for the reference driver's `embed_queries`
    (passage_retrieval.py:32-62): same batching, lower-casing / normalisation and
    tokenizer call, but the encoder's outputs stay on the GPU -- no `.cpu()` per batch (:55) and
    no `.numpy()` at the end (:62).  The CUDA tensor it returns goes straight into
    `Indexer.search_knn`, which searches it where it lies."""
    import torch
    normalize = None
    if getattr(args, "normalize_text", False):
        normalize = importlib.import_module("src.normalize_text").normalize
    model.eval()
    outputs, batch = [], []
    with torch.no_grad():
        for i, question in enumerate(queries):
            if args.lowercase:
                question = question.lower()
            if normalize is not None:
                question = normalize(question)
            batch.append(question)
            if len(batch) == args.per_gpu_batch_size or i == len(queries) - 1:
                enc = tokenizer.batch_encode_plus(batch, return_tensors="pt", max_length=args.question_maxlength,
                                                  padding=True, truncation=True)
                enc = {name: t.cuda() for name, t in enc.items()}
                outputs.append(model(**enc))
                batch = []
    embeddings = torch.cat(outputs, dim=0)
    print(f"Questions embeddings shape: {embeddings.size()}")
    return embeddings


def run_script(script: str, patches=None) -> dict:
    """Runs `script` as __main__ like `python script.py`.  With `patches` ({name: object}) the
    module body is executed first WITHOUT its `if __name__ == "__main__":` blocks, the named
    module-level objects are replaced, and only then the main blocks run -- so a function the
    script defines and calls itself (e.g. `embed_queries`) can be swapped without editing the file."""
    if not patches:
        return runpy.run_path(script, run_name="__main__")
    import ast
    with open(script, "rb") as f:
        tree = ast.parse(f.read(), filename=script)

    def is_main_guard(node) -> bool:
        if not isinstance(node, ast.If) or not isinstance(node.test, ast.Compare):
            return False
        t = node.test
        names = [t.left] + list(t.comparators)
        has_name = any(isinstance(n, ast.Name) and n.id == "__name__" for n in names)
        has_main = any(isinstance(n, ast.Constant) and n.value == "__main__" for n in names)
        return has_name and has_main and len(t.ops) == 1 and isinstance(t.ops[0], ast.Eq)

    guards = [n for n in tree.body if is_main_guard(n)]
    body = ast.Module(body=[n for n in tree.body if n not in guards], type_ignores=[])
    ns = {"__name__": "__main__", "__file__": script, "__builtins__": __builtins__}
    exec(compile(body, script, "exec"), ns)
    for name, obj in patches.items():
        if name not in ns:
            raise AttributeError(f"{script} defines no module-level '{name}' to replace")
        ns[name] = obj
    exec(compile(ast.Module(body=guards, type_ignores=[]), script, "exec"), ns)
    return ns


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
    patches = None
    if os.environ.get("B2IP_DEVICE_QUERIES", "0") not in ("", "0"):
        # keep the query embeddings on the GPU between the encoder and the search (N4)
        patches = {"embed_queries": embed_queries_device}
    run_script(script, patches)


if __name__ == "__main__":
    main()
