#!/usr/bin/env python
"""bench.py -- exact top-k inner-product search throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b2ip|reference]

One "step" = one pass of the hot path over one batch of synthetic queries: all `n_queries`
queries against the whole row-sharded corpus (local tcgen05 scoring + fused filter + fp32
rescore per GPU, then the NCCL all-gather + merge when N > 1).  Default workload = BASELINE
config 3: 21M x 768 fp32 L2-normalised synthetic corpus, 100k queries, k = 100 (fits one B200:
64.5 GB fp32 master + 32.3 GB bf16 shadow).  The corpus is fixed as N grows -> "strong".
`--workload c2|c4|c5` selects the other BASELINE configs for profiles/ (c4: bf16-stored
corpus, k = 1000; c5: batches of `--n-queries` (default 64) queries, k = 10, HBM roofline);
the driver's bench line is always the default.

Prints ONE JSON line on rank 0.  `value` = queries/s with queries resident in HBM;
`e2e` = the same through the reference-facing call with HOST buffers (H2D of the queries and
D2H of scores+ids inside the timed region); `roofline` = the tcgen05 scoring kernel
against the measured bf16 peak; `cpu_baseline` = the faiss-equivalent CPU oracle on the box's
host cores on a bounded sample.  `--impl reference` times that CPU implementation alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # see oracle/flatip_oracle.py::_load
os.environ.setdefault("GOMP_SPINCOUNT", "0")

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "czech-contriever_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "queries/sec top-100 exact IP search, 21M x 768"   # BASELINE.json metric (default workload)
UNIT = "queries/s"
CHUNK = 1 << 18          # rows per generated corpus chunk (seeded by global chunk index)

# BASELINE.json configs.  bound = the roofline that applies to the scoring kernel.
WORKLOADS = {
    "c2": {"n_corpus": 1_000_000, "n_queries": 10_000, "k": 100, "store": "f32", "bound": "tensor"},
    "c3": {"n_corpus": 21_000_000, "n_queries": 100_000, "k": 100, "store": "f32", "bound": "tensor"},
    "c4": {"n_corpus": 21_000_000, "n_queries": 100_000, "k": 1000, "store": "bf16", "bound": "tensor"},
    "c5": {"n_corpus": 21_000_000, "n_queries": 64, "k": 10, "store": "f32", "bound": "hbm"},
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2ip", choices=["b2ip", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--n-corpus", type=int, default=None)
    ap.add_argument("--n-queries", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--store", default=None, choices=["f32", "bf16", "f16"])
    ap.add_argument("--shadow", default=None, choices=["bf16", "f16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=1 << 18)
    ap.add_argument("--cpu-sample-queries", type=int, default=4096)
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    for key in ("n_corpus", "n_queries", "k", "store"):
        if getattr(args, key) is None:
            setattr(args, key, w[key])
    args.bound = w["bound"]
    if args.workload == "c5" and args.steps == 3 and "--steps" not in sys.argv:
        args.steps = 200      # latency per batch over >= 200 batches after warm-up
    return args


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
            "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU every 100 ms via NVML."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def gen_rows(torch, lo, hi, d, seed_base, device):
    """Rows [lo,hi) of the synthetic L2-normalised Gaussian matrix; chunk c (global rows
    [c*CHUNK,(c+1)*CHUNK)) always comes from seed seed_base*1000003+c, whatever the sharding."""
    for c in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):
        c0 = c * CHUNK
        gen = torch.Generator(device=device).manual_seed(seed_base * 1000003 + c)
        x = torch.randn((CHUNK, d), generator=gen, device=device, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        a, b = max(lo, c0) - c0, min(hi, c0 + CHUNK) - c0
        yield max(lo, c0), x[a:b]


def cpu_reference_qps(args, sample_rows_host, sample_q_host, steps, warmup):
    """faiss-IndexFlatIP-equivalent CPU search (oracle/flatip_oracle.c, OpenBLAS sgemm blocks +
    reservoir handler) with all host threads, on the bounded sample; QPS is scaled to the full
    corpus size (cost is linear in rows at fixed nq)."""
    from oracle import flatip_oracle as fo
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fo.search(sample_q_host, sample_rows_host, args.k)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    t = sum(ts) / len(ts)
    scale = args.n_corpus / sample_rows_host.shape[0]
    qps = sample_q_host.shape[0] / (t * scale)
    info = {
        "value": qps, "unit": UNIT, "cores": fo.num_threads(), "kind": "port",
        "sample": (f"{sample_q_host.shape[0]} queries x {sample_rows_host.shape[0]} rows of the same "
                   f"synthetic corpus, {t:.2f} s per pass, scaled x{scale:.1f} to {args.n_corpus} rows; "
                   f"faiss-cpu 1.8.0 not installable offline -> CPU restatement, BLAS: {fo.blas_description()}"),
    }
    return info, t


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (restated, see
    oracle/), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    rng = np.random.default_rng(1234)
    rows = rng.standard_normal((args.cpu_sample_rows, args.d), dtype=np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    q = rng.standard_normal((args.cpu_sample_queries, args.d), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    info, t = cpu_reference_qps(args, rows, q, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": info["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * args.n_queries / info["value"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": info,
        "e2e": {"value": info["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(args):
    m = args.n_corpus / 1e6
    return f"queries/sec top-{args.k} exact IP search, {m:g}M x {args.d}"


def workload_config(args, world):
    op = args.store if args.store != "f32" else (args.shadow or "bf16")
    stored = {"f32": "fp32", "bf16": "bf16-stored", "f16": "fp16-stored"}[args.store]
    return {
        "workload": (f"{args.workload.upper()}: {args.n_corpus} x {args.d} {stored} synthetic L2-normalised corpus, "
                     f"{args.n_queries} queries, k={args.k}"),
        "n_corpus": args.n_corpus, "n_queries": args.n_queries, "k": args.k, "d": args.d,
        "parallelism": f"row-shard x{world} (one process per GPU, NCCL all-gather + merge)",
        "l2": f"inputs exceed L2: every step streams the whole {op} operand copy of the corpus (2*d bytes/row)",
        "coarse": f"tcgen05 kind::f16 {op} operands, fp32 accumulate; fp32 rescore (fp64 accumulate)",
    }


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world_env == 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun (the driver launches us that way itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        sys.exit(subprocess.call(cmd))

    import numpy as np
    import torch
    import torch.distributed as dist
    from b2ip import ShardedIndex, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = world_env
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b2ip needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    N, nq, k, d = args.n_corpus, args.n_queries, args.k, args.d
    lo, hi = shard_bounds(N, world, rank)
    index = ShardedIndex(d, device=local_rank, store=args.store)
    if args.shadow is not None:
        index.engine.set_option("shadow_f16", int(args.shadow == "f16"))
    index.engine.reserve(hi - lo)
    t_ing = time.perf_counter()
    sample_host = None
    for g0, rows in gen_rows(torch, lo, hi, d, 1234, dev):
        index.add_local(rows, g0)
        if rank == 0 and sample_host is None and not args.no_cpu_baseline and world == 1:
            sample_host = rows[:min(args.cpu_sample_rows, rows.shape[0])].cpu().numpy()
    if rank == 0 and sample_host is not None and sample_host.shape[0] < args.cpu_sample_rows:
        extra = [sample_host]
        need = args.cpu_sample_rows - sample_host.shape[0]
        for g0, rows in gen_rows(torch, CHUNK, min(N, CHUNK + need), d, 1234, dev):
            extra.append(rows.cpu().numpy())
        sample_host = np.concatenate(extra)[:args.cpu_sample_rows]
    torch.cuda.synchronize()
    t_ing = time.perf_counter() - t_ing

    gen = torch.Generator(device=dev).manual_seed(4321)
    q_dev = torch.randn((nq, d), generator=gen, device=dev, dtype=torch.float32)
    q_dev /= q_dev.norm(dim=1, keepdim=True)
    q_pin = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
    q_pin.copy_(q_dev)
    D_pin = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
    I_pin = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize()
    index.engine.use_torch_stream()

    # ---------------------------------------------------------------- device-resident timing
    agg = {"coarse_ms": 0.0, "coarse_flops": 0.0, "coarse_launches": 0, "launches": 0,
           "candidates": 0, "rescored": 0, "fallback": 0, "slabs": 0, "refresh_ms": 0.0,
           "finalize_ms": 0.0, "device_ms": 0.0}

    def step_device():
        D, I = index.search(q_dev, k)
        st = index.engine.stats()
        agg["coarse_ms"] += st["coarse_ms"]; agg["coarse_flops"] += st["coarse_flops"]
        agg["coarse_launches"] += st["coarse_launches"]
        # all-gather path: + the NCCL kernel and the merge kernel; the peer-direct exchange's two
        # launches (signal, waiting merge) are already in the library's count
        agg["launches"] += st["total_launches"] + (2 if world > 1 and getattr(index, "exchange_searches", 0) == 0 else 0)
        agg["candidates"] += st["candidates"]; agg["rescored"] += st["rescored"]
        agg["fallback"] += st["fallback_queries"]; agg["slabs"] += st["slabs"]
        agg["refresh_ms"] += st["refresh_ms"]; agg["finalize_ms"] += st["finalize_ms"]
        agg["device_ms"] += st["total_ms"]
        return D, I

    for _ in range(args.warmup):
        step_device()
    for key in agg:
        agg[key] = 0
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        D_last, I_last = step_device()
    e1.record()
    barrier()
    clocks = sampler.finish()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms / args.steps
    value = nq / (ms_per_step / 1e3)

    # ---------------------------------------------------------------- end-to-end timing
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            if world == 1:
                # the reference-facing call: host buffers across the C ABI
                index.engine.search(q_pin.numpy(), k, out=(D_pin.numpy(), I_pin.numpy()))
            else:
                qd = q_pin.to(dev, non_blocking=True)
                D, I = index.search(qd, k)
                D_pin.copy_(D, non_blocking=True)
                I_pin.copy_(I, non_blocking=True)
                torch.cuda.synchronize()
        for _ in range(max(1, args.warmup - 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step_e2e()
        e1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
        e2e = {"value": nq / (ms_e / args.steps / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": world * nq * d * 4, "d2h_bytes_per_step": world * nq * k * 12,
               "ms_per_step": ms_e / args.steps,
               "call": "b2ip_search(mem=HOST) via ctypes" if world == 1 else "ShardedIndex.search + pinned H2D/D2H"}

    # ---------------------------------------------------------------- roofline of the scoring kernel
    peaks, peak_src = load_peaks()
    coarse_ms = max_over_ranks(agg["coarse_ms"])
    flops_rank = agg["coarse_flops"]
    achieved = flops_rank / (agg["coarse_ms"] / 1e3) / 1e12 if agg["coarse_ms"] > 0 else 0.0
    achieved = -max_over_ranks(-achieved)     # slowest rank
    traffic, traffic_note = None, None
    if args.bound == "tensor":
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        try:   # dram bytes of the scoring kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r1_coarse_ncu.json")) as f:
                prof = json.load(f)
            traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
            traffic_note = (f"ncu capture of one launch ({prof['launch']}, {prof['flops']:.3e} FLOP, "
                            f"{prof['duration_ms']:.1f} ms): dram read+write bytes; {prof['source']}")
        except Exception:
            pass
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
            "kernel": "coarse_filter_pair_kernel (tcgen05.mma.cta_group::2 kind::f16, fused threshold filter)",
            "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a multi-second step)",
            "burst_peak": peaks.get("bf16_tflops"),
            "flops_per_launch_avg": flops_rank / max(1, agg["coarse_launches"]),
            "launches": agg["coarse_launches"], "kernel_ms_per_step": coarse_ms / args.steps,
            "kernel_share_of_step": coarse_ms / ms if ms > 0 else None,
        }
    else:
        # small batches: the same kernel streams the 16-bit operand copy of the shard once per
        # batch -> HBM-bound.  Algorithmic bytes = rows * d * 2 (what MUST be read); SURVEY 8d's
        # figure for an fp32-stored corpus (rows * d * 4) is reported next to it.
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_rank = (hi - lo) * d * 2.0 * args.steps
        gbs = bytes_rank / (agg["coarse_ms"] / 1e3) / 1e9 if agg["coarse_ms"] > 0 else 0.0
        gbs = -max_over_ranks(-gbs)
        try:   # dram bytes of the streaming kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r1c_stream_ncu.json")) as f:
                prof = json.load(f)
            traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
            traffic_note = (f"ncu capture of one launch ({prof['launch']}, {prof['algorithmic_bytes']:.4e} "
                            f"algorithmic bytes, {prof['duration_ms']:.3f} ms): dram read+write bytes; {prof['source']}")
        except Exception:
            pass
        roofline = {
            "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
            "traffic": traffic, "traffic_note": traffic_note,
            "kernel": ("coarse_stream_kernel (corpus tile on the MMA's M side, resident queries, TMA-streamed 16-bit rows)"
                       if nq <= 64 else "coarse_filter_kernel<false> (TMA-streamed 16-bit corpus tiles, fused filter)"),
            "peak_source": peak_src + " hbm_gbs (copy bandwidth)",
            "bytes_per_launch_avg": bytes_rank / max(1, agg["coarse_launches"]),
            "launches": agg["coarse_launches"], "kernel_ms_per_step": coarse_ms / args.steps,
            "kernel_share_of_step": coarse_ms / ms if ms > 0 else None,
            "whole_batch_gbs_16bit": (hi - lo) * d * 2.0 / (ms_per_step / 1e3) / 1e9,
            "whole_batch_gbs_fp32_bytes_survey8d": (hi - lo) * d * 4.0 / (ms_per_step / 1e3) / 1e9,
            "ms_per_batch": ms_per_step,
        }

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and sample_host is not None:
        qs = q_dev[:args.cpu_sample_queries].cpu().numpy()
        cpu, _ = cpu_reference_qps(args, sample_host, qs, steps=1, warmup=1)

    launches = int(sum_over_ranks(agg["launches"]))
    # order-independent digest of the last step's results: identical at every N (the sharded
    # search returns the single-GPU answer bit for bit), so the scaling runs check each other
    checksum = {"ids_sum": int(I_last.sum().item()),
                "ids_weighted": int((I_last * torch.arange(1, k + 1, device=dev)).sum().item() % (1 << 61)),
                "scores_sum_f64": float(D_last.double().sum().item())}
    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": f"{args.store if args.store != 'f32' else (args.shadow or 'bf16')} coarse / f32 rescore",
            "data": "synthetic", "config": workload_config(args, world),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "result_checksum": checksum,
            "detail": {"ingest_s": t_ing, "rows_per_gpu": hi - lo,
                       "candidates_per_query_per_step": agg["candidates"] / max(1, args.steps) / nq,
                       "rescored_per_query_per_step": agg["rescored"] / max(1, args.steps) / nq,
                       "fallback_queries": agg["fallback"], "slabs_per_step": agg["slabs"] / max(1, args.steps),
                       "rank0_ms_per_step": {"coarse": agg["coarse_ms"] / max(1, args.steps),
                                             "refresh": agg["refresh_ms"] / max(1, args.steps),
                                             "finalize": agg["finalize_ms"] / max(1, args.steps),
                                             "search_device_total": agg["device_ms"] / max(1, args.steps)},
                       "exchange": ("peer-direct (b2ip_search_exchange)" if getattr(index, "exchange_searches", 0) > 0
                                    else ("nccl all-gather + merge" if world > 1 else "none"))},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
