import json,sys
d=json.loads([l for l in sys.stdin if l.startswith("{")][0])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "dev", round(d["device_ms_per_step"],4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4), d["e2e"]["cuda_graph_builds_in_arm"])
for k in d["kernels"]: print("  ", k["kernel"], round(k["ms_per_launch"]*1e3,1), k["ms_per_launch_cold_l2"] and round(k["ms_per_launch_cold_l2"]*1e3,1))
