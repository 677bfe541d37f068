import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line[:300]); continue
    d=json.loads(line)
    print("grid", d["config"]["grid"], "ms/step %.3f" % d["ms_per_step"], "Gvox/s %.2f" % (d["value"]/1e9), d["phase_ms"], "passes", d["config"]["jacobi_passes_per_step"], "step_roof", d["step_roofline"]["frac"], "nominal", d["step_roofline"]["nominal_bytes_step_formula"]["frac"], "jacobi_roof", d["roofline"]["frac"], "e2e %.2f" % (d["e2e"]["value"]/1e9))
    if "c3" in d:
        c=d["c3"]; print("  c3 256: ms/step %.3f" % c["ms_per_step"], "Gvox/s %.2f" % (c["value"]/1e9), c["phase_ms"], c["step_roofline_frac"])
    if "cpu_baseline" in d: print("  cpu", d["cpu_baseline"])
