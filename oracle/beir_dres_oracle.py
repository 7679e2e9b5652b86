"""CPU restatement of beir 2.0.0 `DenseRetrievalExactSearch.search`
(beir/retrieval/search/dense/exact_search.py), the class the reference instantiates at
src/beir_utils.py:167 and drives through `EvaluateRetrieval.retrieve` (:194).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: beir==2.0.0 (reference environment.yml:132) is a
third-party dependency that is neither vendored nor installable offline; the algorithm below is
restated from the published source: encode queries, sort documents longest first, for each
50k-document chunk `torch.mm` scores (`cos_sim` normalises both sides), NaN -> -1,
`torch.topk(top_k + 1)`, then a per-query `heapq` of size top_k that skips `corpus_id == query_id`.
"""
import heapq
from typing import Dict

import torch


def _cos_sim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a = torch.nn.functional.normalize(a, p=2, dim=1)
    b = torch.nn.functional.normalize(b, p=2, dim=1)
    return torch.mm(a, b.transpose(0, 1))


def _dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return torch.mm(a, b.transpose(0, 1))


def dres_search(model, corpus: Dict[str, Dict[str, str]], queries: Dict[str, str], top_k: int,
                score_function: str, batch_size: int = 128, corpus_chunk_size: int = 50000
                ) -> Dict[str, Dict[str, float]]:
    fn = {"cos_sim": _cos_sim, "dot": _dot}[score_function]
    query_ids = list(queries.keys())
    results = {qid: {} for qid in query_ids}
    qe = torch.as_tensor(model.encode_queries([queries[q] for q in queries], batch_size=batch_size)).float()
    corpus_ids = sorted(corpus, key=lambda k: len(corpus[k].get("title", "") + corpus[k].get("text", "")),
                        reverse=True)
    docs = [corpus[cid] for cid in corpus_ids]
    heaps = {qid: [] for qid in query_ids}
    for start in range(0, len(docs), corpus_chunk_size):
        end = min(start + corpus_chunk_size, len(docs))
        ce = torch.as_tensor(model.encode_corpus(docs[start:end], batch_size=batch_size)).float()
        scores = fn(qe, ce)
        scores[torch.isnan(scores)] = -1
        vals, idx = torch.topk(scores, min(top_k + 1, scores.shape[1]), dim=1, largest=True, sorted=False)
        vals, idx = vals.tolist(), idx.tolist()
        for qi, qid in enumerate(query_ids):
            for sub, score in zip(idx[qi], vals[qi]):
                cid = corpus_ids[start + sub]
                if cid != qid:
                    if len(heaps[qid]) < top_k:
                        heapq.heappush(heaps[qid], (score, cid))
                    else:
                        heapq.heappushpop(heaps[qid], (score, cid))
    for qid in heaps:
        for score, cid in heaps[qid]:
            results[qid][cid] = score
    return results
