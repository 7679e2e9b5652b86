"""Host half of Indexer.search_knn at C3 shape (21M ids, 100k x 100 result rows): numpy
fancy-index + tolist (round 1) vs the C extension csrc/hostmap.c.  CPU only."""
import importlib.util, glob, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = glob.glob(os.path.join(ROOT, "czech-contriever_b200", "lib", "_b2ip_hostmap*.so"))[0]
spec = importlib.util.spec_from_file_location("_b2ip_hostmap", so)
hm = importlib.util.module_from_spec(spec); spec.loader.exec_module(hm)
N, nq, k = int(sys.argv[1]) if len(sys.argv) > 1 else 21_000_000, 100_000, 100
t = time.perf_counter(); ids = [str(i) for i in range(N)]; print(f"make {N} str ids {time.perf_counter()-t:.2f}s")
rng = np.random.default_rng(0)
I = rng.integers(0, N, size=(nq, k), dtype=np.int64)
D = rng.random((nq, k), dtype=np.float32)
t = time.perf_counter(); arr = np.array(ids, dtype=object); t1 = time.perf_counter() - t
t = time.perf_counter(); db = arr[I].tolist(); res = [(db[i], D[i]) for i in range(nq)]; t2 = time.perf_counter() - t
print(f"numpy: object array {t1:.2f}s, map {t2:.3f}s")
del db, res
print("threads", hm.map_threads())
res2 = None
for rep in range(3):
    res2 = None                                   # free the previous result outside the timed call
    t = time.perf_counter(); res2 = hm.map_ids(ids, I, nq, k, list(D)); t3 = time.perf_counter() - t
    print(f"hostmap.map_ids {t3:.3f}s  ({nq*k/t3/1e6:.1f} M ids/s)")
assert res2[5][0] == [ids[i] for i in I[5]] and res2[5][1] is not None and len(res2) == nq
