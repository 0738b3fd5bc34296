import json,sys
d=json.load(open(sys.argv[1]))
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"] if d.get("e2e") else None, d.get("cpu_baseline",{}).get("value"), d["clocks"])
