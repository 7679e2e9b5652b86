"""Latency regime (BASELINE config 5) on N GPUs, one torchrun job: the row-sharded 21M-row corpus is
built once, then option values are alternated between timing passes of bench.py's own
`timed_device_steps` (barrier, CUDA events, max over ranks) -- e.g. the threshold bootstrap on / off.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29512 tools/c5_multi_ab.py --option bootstrap --values 0,1,0,1

Prints one JSON line per (value, batch) on rank 0."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "czech-contriever_b200")]
import torch
import torch.distributed as dist
import bench
from b2ip import ShardedIndex, shard_bounds

ap = argparse.ArgumentParser()
ap.add_argument("--n-corpus", type=int, default=21_000_000)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--batches", default="64,1")
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--option", default="bootstrap")
ap.add_argument("--values", default="0,1,0,1")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
D = bench.Dist(torch, dist, world, dev)
lo, hi = shard_bounds(a.n_corpus, world, rank)
index = ShardedIndex(a.d, device=local, store="f32")
index.engine.reserve(hi - lo)
for g0, rows in bench.gen_rows(torch, lo, hi, a.d, 1234, dev):
    index.add_local(rows, g0)
index.engine.use_torch_stream()
q_all = bench.gen_queries(torch, 64, a.d, dev)
floor_ms = (hi - lo) * a.d * 2 / 6551e9 * 1e3
ref = {}
for v in [int(x) for x in a.values.split(",")]:
    index.engine.set_option(a.option, v)
    for nq in [int(x) for x in a.batches.split(",")]:
        q = q_all[:nq].contiguous()
        ms, agg, last, _ = bench.timed_device_steps(torch, D, index, q, a.k, a.steps, 20, world, light=True)
        Dl, Il, _ = last
        ck = torch.stack([Il.sum(), (Il * torch.arange(1, a.k + 1, device=dev)).sum()])
        if world > 1:
            dist.all_reduce(ck)
        ck = [int(x) for x in ck.tolist()]
        same = ref.setdefault(nq, ck) == ck
        st = index.engine.stats()
        if rank == 0:
            print(json.dumps({"option": a.option, "value": v, "n_gpus": world, "nq": nq, "k": a.k, "rows_per_gpu": hi - lo,
                              "ms_per_batch": round(ms, 4), "floor_ms_16bit": round(floor_ms, 4),
                              "frac_of_floor": round(floor_ms / ms, 4), "sample_rows": st.get("sample_rows", 0),
                              "slabs": st["slabs"], "launches": st["total_launches"],
                              "rank0_coarse_ms": round(agg["coarse_ms"] / a.steps, 4),
                              "rank0_refresh_ms": round(agg["refresh_ms"] / a.steps, 4),
                              "rank0_finalize_ms": round(agg["finalize_ms"] / a.steps, 4),
                              "rank0_device_ms": round(agg["device_ms"] / a.steps, 4),
                              "fallback": int(agg["fallback"]), "checksum_same_as_first": same}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
