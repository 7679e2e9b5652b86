timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2y_tests.log 2>&1; tail -4 gpurun_out/r2y_tests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench1.json 2> gpurun_out/r2y_bench1.err; tail -c 600 gpurun_out/r2y_bench1.err; python tools/show_bench.py gpurun_out/r2y_bench1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
