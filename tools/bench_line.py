"""Print the headline fields of a bench.py JSON line: python tools/bench_line.py file.json [tag]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(sys.argv[2] if len(sys.argv) > 2 else "", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "device", round(d.get("device_ms_per_step", 0), 4),
      "| e2e", round(d["e2e"]["value"], 1), "ms", round(d["e2e"]["ms_per_step"], 4), "| parity", d.get("solution", {}).get("max_rel_state_deviation_vs_cpu_port"))
