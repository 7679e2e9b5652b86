"""Pretty-prints the JSON line of a bench.py run (last line starting with '{' of the file)."""
import json, sys
l = json.loads([x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")][-1])
print(f"N={l['n_gpus']} value {l['value']:.0f} QPS  ms/step {l['ms_per_step']:.2f}  steps {l['steps']}")
for key in ("e2e", "e2e_search_knn"):
    e = l.get(key)
    if e:
        print(f"  {key}: {e['value']:.0f} QPS  ms {e.get('ms_per_step', e.get('ms_per_call', 0)):.1f}  ", {k: v for k, v in e.items() if k in ('matches_device_result', 'matches_engine_result', 'vs_device_step', 'devices')})
r = l["roofline"]
print(f"  roofline {r['achieved']:.1f} {r['unit']} frac {r['frac']:.3f} share {r['kernel_share_of_step']:.3f}  clocks {l['clocks']}")
print("  probe", l["parity_probe"]["pass"], l["parity_probe"]["max_rel_score_err"], "checksum", l["result_checksum"])
print("  detail", {k: v for k, v in l["detail"].items() if k != "rank0_ms_per_step"}, l["detail"]["rank0_ms_per_step"])
if l.get("cpu_baseline"):
    c = l["cpu_baseline"]
    print("  cpu", c.get("value"), c.get("cores"), c.get("seconds_per_step"), [(e["batch"], round(e["ms_per_batch"])) for e in c.get("extra_legs", [])])
for k, v in (l.get("secondary") or {}).items():
    r = v["roofline"]
    print(f"  {k}: {v['value']:.1f} QPS  {v['ms_per_step']:.4f} ms  roof {r['achieved']:.1f} {r['unit']} frac {r['frac']:.3f} whole {r.get('whole_batch_frac_16bit')}  probe {v['parity_probe']['pass']} clk {(v.get('clocks') or {}).get('sm_mhz')} graph {v.get('graph_replay')} resc/q {v['rescored_per_query']:.1f} {v['rank0_ms_per_step']}")
