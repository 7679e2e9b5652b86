"""Seeded randomised differential test: the CUDA path (through the C ABI) against the CPU oracle
over random shapes, k, dimensions, storage types, input dtypes, scales, duplicate rows and
chunked ingest.  Same acceptance as everywhere: identical top-k sets up to ties within 1e-5
relative, scores within 1e-5 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice([64, 128, 256, 384, 768]))
    n = int(rng.integers(1, 60_000))
    nq = int(rng.choice([1, 3, 17, 64, 65, 128, 129, 300]))
    k = int(rng.choice([1, 5, 10, 100, 128, 333, 1000]))
    store = str(rng.choice(["f32", "f32", "f16", "bf16"]))
    scale = float(rng.choice([1.0, 0.05, 30.0]))
    normalize = bool(rng.integers(0, 2))
    dup = bool(rng.integers(0, 4) == 0)
    chunks = int(rng.integers(1, 5))
    return dict(seed=seed, d=d, n=n, nq=nq, k=k, store=store, scale=scale, normalize=normalize, dup=dup,
                chunks=chunks)


@pytest.mark.parametrize("seed", list(range(28)))
def test_random_configuration_matches_oracle(seed):
    import torch
    from oracle import flatip_oracle as fo
    from b2ip import Engine
    c = _case(seed)
    rng = np.random.default_rng(c["seed"])
    x = rng.standard_normal((c["n"], c["d"]), dtype=np.float32)
    q = rng.standard_normal((c["nq"], c["d"]), dtype=np.float32)
    if c["normalize"]:
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-20)
        q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-20)
    x *= c["scale"]
    if c["dup"] and c["n"] > 10:                      # exact ties: some rows repeated
        idx = rng.integers(0, c["n"], size=max(1, c["n"] // 7))
        x[idx] = x[rng.integers(0, c["n"], size=len(idx))]
    # what the index holds after ingest (the oracle sees the same values)
    if c["store"] == "f16":
        held = np.clip(x, -65504, 65504).astype(np.float16).astype(np.float32)
    elif c["store"] == "bf16":
        held = torch.from_numpy(x).bfloat16().float().numpy()
    else:
        held = x
    e = Engine(c["d"], 0, store=c["store"])
    bounds = np.linspace(0, c["n"], c["chunks"] + 1).astype(int)
    for i, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        part = x[a:b]
        e.add(torch.from_numpy(part).cuda() if i % 2 else part)   # host and device buffers
    assert e.ntotal == c["n"]
    D, I = e.search(q, c["k"])
    Do, Io = fo.search(q, held, c["k"])
    assert (np.diff(D, axis=1) <= 0).all(), c
    try:
        fo.compare_topk(D, I, Do, Io, q, held, rtol=RTOL)
    except AssertionError as err:
        raise AssertionError(f"{c}: {err}") from None
    De, Ie = e.search(q, c["k"], mode="exact")
    fo.compare_topk(De, Ie, Do, Io, q, held, rtol=RTOL)
    e.close()
