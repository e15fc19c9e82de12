import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d["config"]["workload"][:60], "%.3e"%d["value"], "%.2f ms"%d["ms_per_step"], {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_step"].items()})
