"""GPU tuning sweep for the scoring kernel: builds one synthetic index, then times searches
under different knob settings (raster group width, L2 eviction hints).
    python tools/tune_coarse.py --n-corpus 4000000 --n-queries 100000 --sweep "gx=32,hint_q=0,hint_x=0;gx=32,hint_q=1,hint_x=2"
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "czech-contriever_b200")]
import torch
from b2ip import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--n-corpus", type=int, default=4_000_000)
ap.add_argument("--n-queries", type=int, default=100_000)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--sweep", default="gx=16")
ap.add_argument("--verbose", type=int, default=0)
ap.add_argument("--store", default="f32")
a = ap.parse_args()
dev = torch.device("cuda", 0)
e = Engine(a.d, 0, store=a.store)
e.reserve(a.n_corpus)
CH = 1 << 18
for c0 in range(0, a.n_corpus, CH):
    g = torch.Generator(device=dev).manual_seed(1234 * 1000003 + c0 // CH)
    x = torch.randn((min(CH, a.n_corpus - c0), a.d), generator=g, device=dev)
    x /= x.norm(dim=1, keepdim=True)
    e.add(x)
g = torch.Generator(device=dev).manual_seed(4321)
q = torch.randn((a.n_queries, a.d), generator=g, device=dev)
q /= q.norm(dim=1, keepdim=True)
e.use_torch_stream()
e.set_option("verbose", a.verbose)
for cfg in a.sweep.split(";"):
    opts = dict(kv.split("=") for kv in cfg.split(",") if kv)
    for name, v in opts.items():
        e.set_option(name, int(v))
    best = None
    for _ in range(a.reps):
        e.search(q, a.k)
        st = e.stats()
        if best is None or st["total_ms"] < best["total_ms"]:
            best = st
    tf = best["coarse_flops"] / best["coarse_ms"] / 1e9
    print(json.dumps({"cfg": cfg, "coarse_ms": round(best["coarse_ms"], 2), "total_ms": round(best["total_ms"], 2),
                      "refresh_ms": round(best["refresh_ms"], 2), "finalize_ms": round(best["finalize_ms"], 2),
                      "tflops": round(tf, 1), "cand_per_q": best["candidates"] / a.n_queries,
                      "rescored_per_q": best["rescored"] / a.n_queries, "slabs": best["slabs"]}), flush=True)
