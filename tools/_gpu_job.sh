K='regex:coarse_|finalize_kernel|refresh_threshold|prep_queries|merge_topk|exchange_|shadow_rows|fill_'
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ac_bench_plain.json 2> gpurun_out/r2ac_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2ac_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ac_bench_ncu.json 2> gpurun_out/r2ac_bench_ncu.err; wc -l gpurun_out/r2ac_bench_launches.csv
export B2IP_GRAPH=0
ncu --set full --import-source on --clock-control none -k regex:refresh_threshold -s 2 -c 1 -f -o gpurun_out/r2ac_refresh_k100_f32 python tools/regimes.py --n-corpus 2625000 --cases 100000:100 --reps 0 > gpurun_out/r2ac_ncu_ref.log 2>&1; tail -1 gpurun_out/r2ac_ncu_ref.log | cut -c1-160
ls -la gpurun_out/r2ac_*
