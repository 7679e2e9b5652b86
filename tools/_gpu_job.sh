python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_ref.json 2> gpurun_out/r2q_ref.err; tail -c 400 gpurun_out/r2q_ref.err; python -c "
import json; l=json.loads(open('gpurun_out/r2q_ref.json').read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['cpu_baseline']['sample'][:200], l['cpu_baseline']['gflops'], l['cpu_baseline']['corpus_to_host_s'])"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench1.json 2> gpurun_out/r2q_bench1.err; tail -c 600 gpurun_out/r2q_bench1.err; python tools/show_bench.py gpurun_out/r2q_bench1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
