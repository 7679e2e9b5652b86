"""BEIR-compatible exact dense search on the B200 engine.

The reference's BEIR path never reaches `Indexer`: `src/beir_utils.py:167-180` builds
`beir.retrieval.search.dense.DenseRetrievalExactSearch(DenseEncoderModel(...), batch_size)` and
`EvaluateRetrieval(dmodel, score_function).retrieve(corpus, queries)` (`:194`) calls
`dmodel.search(corpus, queries, top_k, score_function)`, which in beir 2.0.0 scores 50k-passage
chunks with `torch.mm` on the CPU, takes `torch.topk(top_k+1)` per chunk and merges through a
Python `heapq` per query.  This class keeps that interface (constructor, `search` signature,
`{query_id: {doc_id: score}}` result, longest-document-first encode order, self-match exclusion,
`dot` / `cos_sim`) and replaces the chunked scoring + heap merge with one exact top-(k+1) search
on the GPU.  To use it, pass an instance where the reference constructs beir's class:

    dmodel = b2ip.beir_search.DenseRetrievalExactSearch(DenseEncoderModel(...), batch_size=batch_size)
    retriever = EvaluateRetrieval(dmodel, score_function=score_function)

Differences: NaN scores are dropped instead of being reported as -1 (beir: `cos_scores[isnan] = -1`);
exactly tied scores may keep a different doc id (beir's heap breaks ties by doc-id string).
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _as_numpy(x) -> np.ndarray:
    if type(x).__module__.startswith("torch"):
        x = x.detach().float().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


def _normalize(x: np.ndarray) -> np.ndarray:
    # beir.retrieval.search.dense.util.cos_sim: torch.nn.functional.normalize(p=2, dim=1), eps=1e-12
    n = np.linalg.norm(x, axis=1, keepdims=True)
    return x / np.maximum(n, 1e-12)


class DenseRetrievalExactSearch:
    def __init__(self, model, batch_size: int = 128, corpus_chunk_size: int = 50000, device=None,
                 sharded: bool = None, **kwargs):
        self.model = model                      # encode_queries / encode_corpus, as beir expects
        self.batch_size = batch_size
        self.score_functions = {"cos_sim": "cos_sim", "dot": "dot"}
        self.score_function_desc = {"cos_sim": "Cosine Similarity", "dot": "Dot Product"}
        self.corpus_chunk_size = corpus_chunk_size
        self.show_progress_bar = kwargs.get("show_progress_bar", True)
        self.convert_to_tensor = kwargs.get("convert_to_tensor", True)
        self.device = device
        self.sharded = sharded
        self.results = {}

    def _make_index(self, d: int):
        import torch
        use_shards = self.sharded
        if use_shards is None:
            use_shards = torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1
        dev = self.device if self.device is not None else torch.cuda.current_device()
        if use_shards:
            from .sharded import ShardedIndex
            return ShardedIndex(d, device=dev), True
        from .engine import Engine
        return Engine(d, dev), False

    def search(self, corpus: Dict[str, Dict[str, str]], queries: Dict[str, str], top_k: int,
               score_function: str, return_sorted: bool = False, **kwargs) -> Dict[str, Dict[str, float]]:
        if score_function not in self.score_functions:
            raise ValueError("score function: {} must be either (cos_sim) for cosine similarity or (dot) "
                             "for dot product".format(score_function))
        query_ids = list(queries.keys())
        self.results = {qid: {} for qid in query_ids}
        query_texts = [queries[qid] for qid in queries]
        q = _as_numpy(self.model.encode_queries(
            query_texts, batch_size=self.batch_size, show_progress_bar=self.show_progress_bar,
            convert_to_tensor=self.convert_to_tensor))
        # longest documents first, as beir does (keeps the encoder's batches homogeneous)
        corpus_ids = sorted(corpus, key=lambda c: len(corpus[c].get("title", "") + corpus[c].get("text", "")),
                            reverse=True)
        docs = [corpus[cid] for cid in corpus_ids]
        cos = score_function == "cos_sim"
        if cos:
            q = _normalize(q)
        index, sharded = self._make_index(q.shape[1])
        for start in range(0, len(docs), self.corpus_chunk_size):
            emb = _as_numpy(self.model.encode_corpus(
                docs[start:start + self.corpus_chunk_size], batch_size=self.batch_size,
                show_progress_bar=self.show_progress_bar, convert_to_tensor=self.convert_to_tensor))
            if cos:
                emb = _normalize(emb)
            if sharded:
                index.add_replicated(emb)
            else:
                index.add(emb)
        kk = min(top_k + 1, len(docs))            # +1: room to drop the query's own document
        if kk == 0 or len(query_ids) == 0:
            return self.results
        if sharded:
            import torch
            D, I = index.search(torch.from_numpy(q).cuda(index.engine.device), kk)
            D, I = D.cpu().numpy(), I.cpu().numpy()
        else:
            D, I = index.search(q, kk)
        ids = np.array(corpus_ids, dtype=object)
        for qi, qid in enumerate(query_ids):
            res = self.results[qid]
            for score, row in zip(D[qi].tolist(), I[qi].tolist()):
                if row < 0:
                    break
                cid = ids[row]
                if cid != qid:
                    if len(res) < top_k:
                        res[cid] = score
        return self.results
