import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line[:300]); continue
    d=json.loads(line)
    print("grid", d["config"]["grid"], "ms/step %.3f" % d["ms_per_step"], "Gvox/s %.2f" % (d["value"]/1e9), d["phase_ms"], "passes", d["config"]["jacobi_passes_per_step"], "step_roof", d["step_roofline"]["frac"], "nominal", d["step_roofline"]["nominal_bytes_step_formula"]["frac"], "jacobi_roof", d["roofline"]["frac"], "e2e %.2f" % (d["e2e"]["value"]/1e9))
    for key in ("c3", "c2"):
        if key in d:
            c=d[key]; print("  %s %s: ms/step %.3f" % (key, c["workload"][3:8], c["ms_per_step"]), "Gvox/s %.2f" % (c["value"]/1e9), c["phase_ms"], c["step_roofline_frac"])
    if "cpu_baseline" in d: print("  cpu", d["cpu_baseline"])
    if "e2e_export" in d: print("  e2e with the colour field copied out every step: %.2f Gvox/s" % (d["e2e_export"]["value"]/1e9))
    ex=d.get("experiments")
    if ex:
        print("  experiments (%s s, exit %s):" % (ex.get("seconds"), ex.get("exit")), ex.get("error", ""))
        base={}
        for r in ex.get("results", []):
            if r.get("variant")=="default": base[r["grid"]]=r
        for r in ex.get("results", []):
            g=r.get("grid"); b=base.get(g)
            if "error" in r or "skipped" in r: print("    %-12s %-22s %s" % (g, r.get("variant"), r.get("error", r.get("skipped")))); continue
            if r.get("variant")=="light_map_pass": print("    %-12s %s" % (g, {k:v for k,v in r.items() if k not in ("grid","variant")})); continue
            rel=(" (%+.1f %% vs default)" % (100*(r["ms_per_step"]/b["ms_per_step"]-1))) if b and r is not b and "ms_per_step" in r else ""
            print("    %-12s %-22s %.4f ms/step%s  jacobi %s advect %s%s" % (g, r.get("variant"), r.get("ms_per_step", float("nan")), rel, r.get("jacobi_ms"), r.get("advect_ms"),
                  "" if not r.get("mismatched_elements_vs_default") else "  MISMATCH %d" % r["mismatched_elements_vs_default"]))
