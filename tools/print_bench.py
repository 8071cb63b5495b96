"""Prints the headline fields of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.2f %s | e2e %.2f | ms/step %.1f | latency %.1f ms | launches %s" % (d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"],
                                                                                d.get("latency_ms_per_pair", 0), d.get("gpu_launches")))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "traffic")} if d.get("roofline") else None)
for k in ("roofline_pyramid", "roofline_nn"):
    if d.get(k):
        print(k, round(d[k]["achieved"], 1), d[k]["unit"], "frac %.3f" % d[k]["frac"])
print("cpu_baseline", d.get("cpu_baseline") and (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"]))
print("clocks", d.get("clocks"))
