timeout 300 python tools/c5_latency.py --n-corpus 21000000 --iters 100 --fused 0,1,2,0,1,2 --stages 8 > gpurun_out/r2v_c5_21M.log 2>&1; cut -c1-250 gpurun_out/r2v_c5_21M.log
export B2IP_GRAPH=0
for f in 0 1; do
B2IP_STREAM_FUSED=$f timeout 300 ncu --set full --import-source on --clock-control none -k regex:coarse_stream -c 6 -f -o gpurun_out/r2v_stream_fused$f python tools/regimes.py --n-corpus 21000000 --cases 64:10 --reps 0 > gpurun_out/r2v_ncu_$f.log 2>&1; tail -1 gpurun_out/r2v_ncu_$f.log | cut -c1-200
done
ls -la gpurun_out/r2v_*.ncu-rep
