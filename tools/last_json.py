"""Prints selected fields of the last JSON line of a bench.py output file."""
import json, sys
for path in sys.argv[1:]:
    line = [l for l in open(path) if l.startswith("{")][-1]
    l = json.loads(line)
    e = l.get("e2e") or {}
    r = l.get("roofline") or {}
    print(path, "| n_gpus", l["n_gpus"], "| ms/step", round(l["ms_per_step"], 4), "| value", round(l["value"], 1),
          "| e2e ms", round(e.get("ms_per_step", 0), 4), "| kernel frac", round(r.get("frac", 0), 4),
          "| share", round(r.get("kernel_share_of_step") or 0, 3), "|", l.get("result_checksum", {}).get("ids_sum"), "|", l.get("detail", {}).get("rank0_ms_per_step"),
          l.get("detail", {}).get("exchange"))
