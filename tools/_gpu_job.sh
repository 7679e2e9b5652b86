python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2d_parity.log 2>&1; tail -5 gpurun_out/r2d_parity.log
for cfg in "1 1" "1 0" "0 1"; do set -- $cfg; echo "== graph=$1 timing=$2"
  B2IP_GRAPH=$1 B2IP_GRAPH_TIMING=$2 python bench.py --workload c5 --n-corpus 2625000 --steps 1000 --warmup 50 --no-cpu-baseline --no-search-knn --no-e2e > gpurun_out/r2d_c5_g$1t$2.json 2>gpurun_out/r2d_c5.err
  python tools/show_bench.py gpurun_out/r2d_c5_g$1t$2.json | head -4; done
for nq in 1 16; do B2IP_GRAPH=1 python bench.py --workload c5 --n-corpus 2625000 --n-queries $nq --steps 1000 --warmup 50 --no-cpu-baseline --no-search-knn --no-e2e > gpurun_out/r2d_c5_nq$nq.json 2>>gpurun_out/r2d_c5.err
  python tools/show_bench.py gpurun_out/r2d_c5_nq$nq.json | head -3; done
tail -5 gpurun_out/r2d_c5.err
