"""Prints the essentials of bench.py JSON lines (file argument or stdin)."""
import json
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for line in src:
    line = line.strip()
    if not line.startswith("{"):
        print(line[:300])
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("reference arm: %.4f Gvox/s" % (d["value"] / 1e9), d["cpu_baseline"])
        continue
    print("grid", d["config"]["grid"], "N", d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "Gvox/s %.2f" % (d["value"] / 1e9),
          d["phase_ms"], "passes", d["config"]["jacobi_passes_per_step"], "step_roof", d["step_roofline"]["frac"],
          "dominant", d["roofline"]["kernel"], d["roofline"]["frac"], "e2e %.2f" % (d["e2e"]["value"] / 1e9), "checksum", d["state_checksum"])
    print("   phase fracs", {k: v["frac"] for k, v in d["phase_roofline"].items()}, "bricks", d["config"]["bricks_relaxed_per_step"], d["config"]["bricks_copied_per_step"], "clocks", d["clocks"])
    if d.get("per_rank"):
        for r, pr in enumerate(d["per_rank"]):
            print("   rank %d" % r, pr)
    for key in ("c3", "c2", "c4_strong"):
        if key in d:
            c = d[key]
            print("  %s %s: ms/step %.3f" % (key, c["workload"][:12], c["ms_per_step"]), "Gvox/s %.2f" % (c["value"] / 1e9), c["phase_ms"],
                  c.get("step_roofline", {}).get("frac"), {k: v["frac"] for k, v in c.get("phase_roofline", {}).items()}, c["state_checksum"])
    if "cpu_baseline" in d:
        print("  cpu", d["cpu_baseline"])
    if "e2e_export" in d:
        print("  e2e with the colour field copied out every step: %.2f Gvox/s" % (d["e2e_export"]["value"] / 1e9))
