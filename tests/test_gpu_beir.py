"""N1: the BEIR-compatible searcher against the restatement of beir's own
DenseRetrievalExactSearch (oracle/beir_dres_oracle.py), same fake encoder on both sides."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class HashEncoder:
    """Deterministic stand-in for DenseEncoderModel (src/beir_utils.py:24-133): text -> vector."""

    def __init__(self, d=128):
        self.d = d

    def _vec(self, text):
        seed = int.from_bytes(hashlib.sha1(text.encode()).digest()[:8], "little")
        return np.random.default_rng(seed).standard_normal(self.d).astype(np.float32)

    def encode_queries(self, queries, batch_size, **kw):
        return np.stack([self._vec(q) for q in queries])

    def encode_corpus(self, corpus, batch_size, **kw):
        return np.stack([self._vec(c.get("title", "") + " " + c["text"]) for c in corpus])


def _data(n_docs, n_q):
    rng = np.random.default_rng(0)
    corpus = {f"d{i}": {"title": f"t{i}", "text": "w " * int(rng.integers(1, 40)) + str(i)} for i in range(n_docs)}
    queries = {f"q{i}": f"question {i}" for i in range(n_q)}
    # a query whose id equals a document id: that document must be excluded for it
    queries["d7"] = "t7 " + corpus["d7"]["text"]
    return corpus, queries


@pytest.mark.parametrize("score_function", ["dot", "cos_sim"])
@pytest.mark.parametrize("top_k", [10, 1000])
def test_matches_beir_restatement(score_function, top_k):
    from b2ip.beir_search import DenseRetrievalExactSearch
    from oracle.beir_dres_oracle import dres_search
    corpus, queries = _data(6000, 40)
    enc = HashEncoder()
    got = DenseRetrievalExactSearch(enc, batch_size=64, corpus_chunk_size=2500).search(
        corpus, queries, top_k, score_function)
    want = dres_search(enc, corpus, queries, top_k, score_function, corpus_chunk_size=2500)
    assert got.keys() == want.keys()
    for qid in want:
        assert len(got[qid]) == len(want[qid]) == min(top_k, len(corpus) - (qid in corpus))
        assert qid not in got[qid]
        assert got[qid].keys() == want[qid].keys(), qid     # random data: no exact ties
        for cid, s in want[qid].items():
            assert abs(got[qid][cid] - s) <= 1e-5 * max(abs(s), 1e-2 * (1.0 if score_function == "cos_sim" else 128.0))


def test_rejects_unknown_score_function():
    from b2ip.beir_search import DenseRetrievalExactSearch
    with pytest.raises(ValueError):
        DenseRetrievalExactSearch(HashEncoder()).search({}, {}, 10, "l2")
