"""Turns ncu outputs brought back in gpurun_out/ into the small committed summaries under profiles/.

    python tools/summarize_profiles.py launches <launches.csv> <out.txt>     # per-launch device times + shares
    python tools/summarize_profiles.py full <report.ncu-rep> <out.json>      # key metrics of every kernel in a report
    python tools/summarize_profiles.py traffic <traffic.csv> <NXxNYxNZ>      # DRAM bytes per launch -> profiles/traffic.json
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def short(name):
    name = name.split("(")[0]
    return name.replace("fxb::", "").replace("<unnamed>::", "").replace("void ", "").strip()[-60:]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
    per = {}
    for r in rows[1:]:
        if r[mi] == "gpu__time_duration.sum":
            per[int(r[ii])] = (short(r[ki]), float(r[vi].replace(",", "")) / 1000.0)
    total = sum(v for _, v in per.values())
    agg = {}
    for name, us in per.values():
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        f.write("# total %.1f us over %d launches\n# kernel | launches | total us | share\n" % (total, len(per)))
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-62s %4d %10.1f %6.1f%%\n" % (name, n, us, 100 * us / total))
        f.write("\n# per launch (id, kernel, us)\n")
        for i in sorted(per):
            f.write("%4d %-62s %10.1f\n" % (i, per[i][0], per[i][1]))
    print(open(out).read()[:1500])


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[ki])}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k + (" [" + units[i] + "]" if units[i] else "")] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res[:2], indent=1)[:2500])


PHASES = (("advect", "advect_kernel"), ("divergence", "divergence_quad_kernel"), ("gradient", "gradient_quad_kernel"),
          ("jacobi", "jacobi_"))  # jacobi: the marching first pass + the resident passes + the settle copy


def traffic(path, grid):
    """ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the launches of whole steps: DRAM bytes per launch
    of each phase's kernel(s), averaged over the launches captured (bench.py's roofline.traffic reads the result)."""
    import os
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Metric Unit"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[1:]:
        if r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            e = per.setdefault(int(r[ii]), [short(r[ki]), 0.0])
            e[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    try:
        doc = json.load(open(out_path))
    except Exception:
        doc = {}
    entry = {}
    for phase, prefix in PHASES:
        sel = [b for n, b in per.values() if prefix in n and "flip" not in n]
        if sel:
            entry[phase] = {"dram_bytes_per_launch": sum(sel) / len(sel), "launches_captured": len(sel),
                            "dram_bytes_total": sum(sel),
                            "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                                      "every launch of %d step(s) after 100 spin-up steps (%s)" % (
                                          max(1, len([1 for n, _ in per.values() if "advect_kernel" in n])),
                                          os.path.basename(path))}
    doc[grid] = entry
    json.dump(doc, open(out_path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
