"""torchrun worker for tests/test_gpu_multi.py: row-sharded search over NCCL must equal the
single-GPU search bit for bit (same kernels, merge keeps score desc / lower global row)."""
import os
import sys

os.environ.setdefault("B2IP_EXCHANGE_TIMEOUT_S", "60")     # a lost flag fails the test in a minute, not ten

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
whole = Engine(d, local)
whole.add(x)
for k in (10, 100, 1000):
    Dw, Iw = whole.search(qd, k)
    for peer in (True, False):          # fused peer-direct exchange / NCCL all-gather + merge
        idx = ShardedIndex(d, device=local, peer_exchange=peer)
        lo, hi = shard_bounds(n, world, rank)
        idx.add_local(x[lo:hi], lo)
        D, I = idx.search(qd, k)
        assert torch.equal(I, Iw), f"k={k} peer={peer}: sharded ids differ from single-GPU ids"
        assert torch.equal(D, Dw), f"k={k} peer={peer}: sharded scores differ"
        assert idx.exchange_searches == (1 if peer else 0)
        # round-robin chunk ingest (several row segments per shard) gives the same answer too
        idx2 = ShardedIndex(d, device=local, peer_exchange=peer)
        for a in range(0, n, 7000):
            idx2.add_replicated(x[a:a + 7000])
        D2, I2 = idx2.search(qd, k)
        assert torch.equal(I2, Iw) and torch.equal(D2, Dw), f"k={k} peer={peer}: replicated ingest differs"

# owner mode + global threshold round: every rank holds the global answer of the queries it owns;
# with the threshold round each rank rescores about k/world rows instead of k + a window
nq_big = 2500                                        # > 2048: the threshold round is used from here on
q_big = torch.from_numpy(synth(nq_big, d, 777)).cuda()
for k in (10, 100, 1000):
    for queries, nqs in ((qd, nq), (q_big, nq_big)):
        Dw, Iw = whole.search(queries, k)
        for gthr in (True, False):
            idx = ShardedIndex(d, device=local)
            idx.global_threshold = gthr
            lo, hi = shard_bounds(n, world, rank)
            idx.add_local(x[lo:hi], lo)
            D, I, (qlo, qhi) = idx.search_owned(queries, k)
            assert (qlo, qhi) == shard_bounds(nqs, world, rank) and D.shape == (qhi - qlo, k)
            assert torch.equal(I, Iw[qlo:qhi]) and torch.equal(D, Dw[qlo:qhi]), f"k={k} gthr={gthr}: owned slice differs"
            st = idx.engine.stats()
            assert st["bound_violations"] == 0 and st["max_err_over_eps"] < 1.0
            resc = st["rescored"] / nqs
            if gthr:
                assert idx.exchange_searches == 1
                resc_with = resc
            elif nqs > 2048 and k >= 100:
                assert resc_with < resc, (k, resc_with, resc)    # the global bound prunes the rescore
            else:
                assert resc_with <= resc, (k, resc_with, resc)
            D2, I2 = idx.search(queries, k)                   # the all-ranks form over the same buffers
            assert torch.equal(I2, Iw) and torch.equal(D2, Dw)
    # host buffers in, this rank's slice out (the e2e call of bench.py at N > 1)
    Dh = np.full((nq, k), np.nan, np.float32); Ih = np.full((nq, k), -7, np.int64)
    for qq in (q, q.astype(np.float16)):
        Dref, Iref = whole.search(torch.from_numpy(qq.astype(np.float32)).cuda(), k)
        _, _, (qlo, qhi) = idx.search_host(qq, k, out=(Dh, Ih))
        assert np.array_equal(Ih[qlo:qhi], Iref[qlo:qhi].cpu().numpy()) and np.array_equal(Dh[qlo:qhi], Dref[qlo:qhi].cpu().numpy())
# fewer queries than ranks (a rank that owns nothing), and an empty shard, in owner mode
idx_o = ShardedIndex(d, device=local)
idx_e = ShardedIndex(d, device=local)
lo, hi = shard_bounds(n, world, rank)
idx_o.add_local(x[lo:hi], lo)
if rank == 0:
    idx_e.add_local(x, 0)
for nqs in (1, max(1, world - 1), world, 3 * world + 1, 500):
    Dw, Iw = whole.search(qd[:nqs].contiguous(), 10)
    for index in (idx_o, idx_e):
        D, I, (qlo, qhi) = index.search_owned(qd[:nqs].contiguous(), 10)
        assert torch.equal(I, Iw[qlo:qhi]) and torch.equal(D, Dw[qlo:qhi]), f"owned search nq={nqs} differs"

# peer-direct exchange over many searches: both parities, buffers regrown, batch sizes of the
# latency regime, and a rank with an EMPTY shard
idx = ShardedIndex(d, device=local)
idx3 = ShardedIndex(d, device=local)
lo, hi = shard_bounds(n, world, rank)
idx.add_local(x[lo:hi], lo)
if rank == 0:
    idx3.add_local(x, 0)                # every row on rank 0, nothing anywhere else
for it, nqs in enumerate([1, 7, 64, 64, 500, 3, 128, 129, 500, 1]):
    Dw, Iw = whole.search(qd[:nqs].contiguous(), 10)
    for index in (idx, idx3):
        D, I = index.search(qd[:nqs].contiguous(), 10)
        assert torch.equal(I, Iw) and torch.equal(D, Dw), f"exchange search {it} (nq={nqs}) differs"
assert idx.exchange_searches == 10 and idx3.exchange_searches == 10

# adversarial corpus (ascending scores): candidate lists overflow on the rank that holds the tail;
# every rank sees the same non-zero status and repeats the search through the all-gather path
u = synth(1, 128, 3)[0]
xa = (u[None, :] * (1.0 + np.arange(60_000, dtype=np.float32)[:, None] * 1e-5)).astype(np.float32)
qa = torch.from_numpy(np.tile(u[None, :], (8, 1)) * np.linspace(0.5, 1.5, 8, dtype=np.float32)[:, None]).cuda()
ia = ShardedIndex(128, device=local)
lo, hi = shard_bounds(60_000, world, rank)
ia.add_local(xa[lo:hi], lo)
Da, Ia = ia.search(qa, 100)
assert ia.exchange_searches == 0, "overflow must send every rank to the all-gather path"
assert torch.equal(Ia[0].cpu(), torch.arange(59_999, 59_899, -1)), Ia[0][:5]
if rank == 0:
    # the single-process driver of the same shards (b2ip.multi): one engine per visible GPU
    from b2ip import MultiGpuEngine
    m = MultiGpuEngine(d)
    assert len(m.engines) >= world
    m.add(x[:25_000]); m.add(x[25_000:25_003]); m.add(x[25_003:])
    assert max(mp.n_local for mp in m.maps) - min(mp.n_local for mp in m.maps) <= 1
    for k in (10, 100):
        Dw, Iw = whole.search(q, k)
        Dm, Im = m.search(q, k)
        assert np.array_equal(Im, Iw) and np.array_equal(Dm, Dw), f"k={k}: MultiGpuEngine differs"
        Dm, Im = m.search(q.astype(np.float16), k)        # float16 queries are widened on the device
        Dh, Ih = whole.search(q.astype(np.float16), k)
        assert np.array_equal(Im, Ih) and np.array_equal(Dm, Dh)
        Dt, It = m.search(qd, k)                           # device queries in, tensors out
        assert torch.equal(It.cpu(), torch.from_numpy(Iw)) and torch.equal(Dt.cpu(), torch.from_numpy(Dw))
    assert m.exchange_searches == 6, m.exchange_searches   # the fused exchange, not the copy-and-merge path
    for nqs in (1, 3):
        Dw, Iw = whole.search(q[:nqs], 10)
        Dm, Im = m.search(q[:nqs], 10)
        assert np.array_equal(Im, Iw) and np.array_equal(Dm, Dw)
    m.close()
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank} ok", flush=True)
