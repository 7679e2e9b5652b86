python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q > gpurun_out/r2i_parity.log 2>&1; tail -5 gpurun_out/r2i_parity.log
python tools/regimes.py --n-corpus 2625000 --cases 100000:100,100000:1000 --reps 2 > gpurun_out/r2i_regimes_f32.log 2>&1
python tools/regimes.py --n-corpus 2625000 --store bf16 --cases 100000:100,100000:1000 --reps 2 > gpurun_out/r2i_regimes_bf16.log 2>&1
cat gpurun_out/r2i_regimes_f32.log gpurun_out/r2i_regimes_bf16.log | cut -c1-330
