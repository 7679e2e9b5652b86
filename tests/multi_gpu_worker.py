"""torchrun worker for tests/test_gpu_multi.py: row-sharded search over NCCL must equal the
single-GPU search bit for bit (same kernels, merge keeps score desc / lower global row)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "czech-contriever_b200"), os.path.join(ROOT, "tests")]
from b2ip import Engine, ShardedIndex, shard_bounds  # noqa: E402
from helpers import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, nq, d = 60_000, 500, 768
x, q = synth(n, d, 1234), synth(nq, d, 4321)
x[40_000:40_050] = x[100:150]                       # exact ties that straddle shards
qd = torch.from_numpy(q).cuda()
for k in (10, 100, 1000):
    idx = ShardedIndex(d, device=local)
    lo, hi = shard_bounds(n, world, rank)
    idx.add_local(x[lo:hi], lo)
    D, I = idx.search(qd, k)
    if rank == 0:
        whole = Engine(d, local)
        whole.add(x)
        Dw, Iw = whole.search(qd, k)
        assert torch.equal(I, Iw), f"k={k}: sharded ids differ from single-GPU ids"
        assert torch.equal(D, Dw), f"k={k}: sharded scores differ"
    # round-robin chunk ingest (multi-segment id mapping) gives the same answer too
    idx2 = ShardedIndex(d, device=local)
    for a in range(0, n, 7000):
        idx2.add_replicated(x[a:a + 7000])
    D2, I2 = idx2.search(qd, k)
    assert torch.equal(I2, I) and torch.equal(D2, D), f"k={k}: replicated ingest differs"
if rank == 0:
    # the single-process driver of the same shards (b2ip.multi): one engine per visible GPU
    from b2ip import MultiGpuEngine
    m = MultiGpuEngine(d)
    assert len(m.engines) >= world
    m.add(x[:25_000]); m.add(x[25_000:])
    whole = Engine(d, local)
    whole.add(x)
    for k in (10, 100):
        Dw, Iw = whole.search(q, k)
        Dm, Im = m.search(q, k)
        assert np.array_equal(Im, Iw) and np.array_equal(Dm, Dw), f"k={k}: MultiGpuEngine differs"
    m.close()
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank} ok", flush=True)
