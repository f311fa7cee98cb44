import json,sys
d=json.load(open(sys.argv[1]))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["ipm_iterations_mean"], d["config"]["status_optimal_frac"], d["gpu_launches"])
for k,v in d["kernels"].items(): print(k, round(v["avg_us"],1), v["launches"])
