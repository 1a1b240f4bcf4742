import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line); continue
    d=json.loads(line)
    print("ms/step %.1f  kernels %s  value %.3e  elbo %.3f launches %s  roofline frac %s e2e %s cpu %s" % (d["ms_per_step"], d["kernel_ms"], d["value"], d["elbo_last"], d["gpu_launches"], d["roofline"] and round(d["roofline"]["frac"],4), d.get("e2e") and d["e2e"].get("value"), d.get("cpu_baseline") and d["cpu_baseline"]["value"]))
