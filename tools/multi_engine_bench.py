"""C3 through the SINGLE-PROCESS multi-GPU engine (b2ip.multi.MultiGpuEngine, what
`Indexer(device="all")` uses): every visible GPU driven by host threads of one process, P2P
gather + merge kernel.  Compare with `bench.py --gpus N` (one process per GPU, NCCL).
    python tools/multi_engine_bench.py [--n-corpus 21000000] [--n-queries 100000] [--k 100]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "czech-contriever_b200")]
import numpy as np, torch
from b2ip import MultiGpuEngine
from bench import gen_rows

ap = argparse.ArgumentParser()
ap.add_argument("--n-corpus", type=int, default=21_000_000)
ap.add_argument("--n-queries", type=int, default=100_000)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
m = MultiGpuEngine(a.d)
m.reserve(a.n_corpus)
t0 = time.perf_counter()
for g0, rows in gen_rows(torch, 0, a.n_corpus, a.d, 1234, dev):
    m.add(rows)
torch.cuda.synchronize()
t_ing = time.perf_counter() - t0
gen = torch.Generator(device=dev).manual_seed(4321)
q = torch.randn((a.n_queries, a.d), generator=gen, device=dev)
q /= q.norm(dim=1, keepdim=True)
qh = q.cpu().numpy()
for name, arg in (("device queries", q), ("host (pageable numpy) queries + results", qh)):
    best = 1e9
    for _ in range(a.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        D, I = m.search(arg, a.k)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    st = m.stats()
    print(json.dumps({"path": name, "gpus": len(m.engines), "n": a.n_corpus, "nq": a.n_queries, "k": a.k,
                      "wall_ms": round(best * 1e3, 2), "qps": round(a.n_queries / best, 1),
                      "coarse_ms_max": round(st["coarse_ms"], 2), "ingest_s": round(t_ing, 2),
                      "ids_sum": int(np.asarray(I.cpu() if hasattr(I, "cpu") else I).sum())}), flush=True)
