"""Host->index ingest rate of Engine.add (reference: index.add at src/index.py:30)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "czech-contriever_b200")]
import numpy as np, torch
from b2ip import Engine
n, d = 1_000_000, 768
x32 = np.random.default_rng(0).standard_normal((n, d), dtype=np.float32)
x16 = x32.astype(np.float16)
for name, arr in (("fp32 pageable", x32), ("fp16 pageable", x16)):
    for rep in range(2):
        e = Engine(d, 0); e.reserve(n)
        t = time.perf_counter(); e.add(arr); dt = time.perf_counter() - t
        print(f"{name}: {n} rows in {dt*1e3:.1f} ms = {arr.nbytes/dt/1e9:.2f} GB/s host bytes, {n/dt/1e6:.2f} M rows/s", flush=True)
        e.close()
pin = torch.from_numpy(x32).pin_memory()
e = Engine(d, 0); e.reserve(n)
t = time.perf_counter(); e.add(pin.numpy()); dt = time.perf_counter() - t
print(f"fp32 pinned: {dt*1e3:.1f} ms = {x32.nbytes/dt/1e9:.2f} GB/s")
