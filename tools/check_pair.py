"""Quick GPU check of the CTA-pair scoring kernel against the 1-CTA kernel and the exact path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "czech-contriever_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from b2ip import Engine
from helpers import synth
for n, nq, k in [(20000, 300, 100), (70001, 1000, 10), (5000, 129, 100), (300000, 4096, 100)]:
    x, q = synth(n, 768, 1), synth(nq, 768, 2)
    e = Engine(768, 0); e.add(x)
    e.set_option("pair", 1); Dp, Ip = e.search(q, k, mode="tensor"); sp = e.stats()
    e.set_option("pair", 0); D1, I1 = e.search(q, k, mode="tensor"); s1 = e.stats()
    De, Ie = e.search(q[:64], k, mode="exact")
    ok = np.array_equal(Ip, I1) and np.array_equal(Dp, D1) and np.array_equal(Ip[:64], Ie)
    print(n, nq, k, "pair==single==exact:", ok, "cand", sp["candidates"], s1["candidates"],
          "ms", round(sp["coarse_ms"], 3), round(s1["coarse_ms"], 3), flush=True)
    assert ok
print("pair kernel ok")
