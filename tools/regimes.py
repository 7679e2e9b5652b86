"""Times the search in every BASELINE regime on ONE GPU holding one shard-sized corpus:
large batch k=100 (C3 per-GPU shard), k=1000 (C4), small batches k=10 (C5), and C2 (1M x 10k)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "czech-contriever_b200")]
import torch
from b2ip import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--n-corpus", type=int, default=2_625_000)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--cases", default="100000:100,100000:1000,10000:100,1:10,4:10,16:10,64:10,256:10,2048:100")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--store", default="f32")
ap.add_argument("--shadow", default=None, help="bf16 | f16 (store=f32 only)")
ap.add_argument("--verbose", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)


def build(n):
    e = Engine(a.d, 0, store=a.store, shadow=a.shadow)
    e.reserve(n)
    CH = 1 << 18
    for c0 in range(0, n, CH):
        g = torch.Generator(device=dev).manual_seed(1234 * 1000003 + c0 // CH)
        x = torch.randn((min(CH, n - c0), a.d), generator=g, device=dev)
        x /= x.norm(dim=1, keepdim=True)
        e.add(x)
    e.use_torch_stream()
    e.set_option("verbose", a.verbose)
    return e


e = build(a.n_corpus)
g = torch.Generator(device=dev).manual_seed(4321)
qall = torch.randn((100_000, a.d), generator=g, device=dev)
qall /= qall.norm(dim=1, keepdim=True)
hbm_ms = a.n_corpus * a.d * 2 / 6551e9 * 1e3
for case in a.cases.split(","):
    nq, k = map(int, case.split(":"))
    q = qall[:nq].contiguous()
    best_wall, best = 1e9, None
    for _ in range(a.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e.search(q, k)
        torch.cuda.synchronize()
        w = (time.perf_counter() - t0) * 1e3
        if w < best_wall:
            best_wall, best = w, e.stats()
    flops = 2.0 * nq * a.n_corpus * a.d
    print(json.dumps({"store": a.store, "shadow": a.shadow, "n": a.n_corpus, "nq": nq, "k": k, "wall_ms": round(best_wall, 3),
                      "dev_ms": round(best["total_ms"], 3), "coarse_ms": round(best["coarse_ms"], 3), "refresh_ms": round(best["refresh_ms"], 3), "finalize_ms": round(best["finalize_ms"], 3),
                      "tflops_wall": round(flops / best_wall / 1e9, 1), "hbm_floor_ms_bf16": round(hbm_ms, 3),
                      "qps": round(nq / best_wall * 1e3, 1), "slabs": best["slabs"], "launches": best["total_launches"],
                      "cand_per_q": round(best["candidates"] / nq, 1), "resc_per_q": round(best["rescored"] / nq, 1),
                      "fallback": best["fallback_queries"]}), flush=True)
