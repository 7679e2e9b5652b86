python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; tail -6 gpurun_out/r2j_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 1500 gpurun_out/r2j_bench.err; python tools/show_bench.py gpurun_out/r2j_bench.json
