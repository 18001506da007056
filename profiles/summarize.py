"""Turns an `ncu -i <rep> --page raw --csv` dump of the path's kernels into the per-launch summary (JSON + a markdown
table) and refreshes profiles/fit_traffic.json, which bench.py reads for roofline.traffic.

    python profiles/summarize.py profiles/r01g_path_full_raw.csv r01g [points_per_launch_group]
"""
import csv, json, re, sys
from pathlib import Path

raw = Path(sys.argv[1]); tag = sys.argv[2]
points = int(sys.argv[3]) if len(sys.argv) > 3 else 64 * 120000
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def to_mb(v, unit):
    v = float(v)
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[unit]


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)


out = []
for r in data:
    name = re.sub(r"^void |rpw::|\(.*$", "", r[col["Kernel Name"]])
    rec = {"kernel": name}
    for k in keep:
        if k in col:
            rec[k] = r[col[k]]; rec[k + "__unit"] = units[col[k]]
    out.append(rec)
Path(raw.parent / f"{tag}_path_full_summary.json").write_text(json.dumps(out, indent=1))

print("| kernel | time us | DRAM read MB | DRAM write MB | DRAM % of peak | SM throughput % | warps active % | issue active % | regs | grid x block | warp instr M |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
tot = {"fit_r": 0.0, "fit_w": 0.0}
per = {}
for rec in out:
    g = lambda k: rec.get(k, "0")
    rd = to_mb(g("dram__bytes_read.sum"), g("dram__bytes_read.sum__unit"))
    wr = to_mb(g("dram__bytes_write.sum"), g("dram__bytes_write.sum__unit"))
    t = to_us(g("gpu__time_duration.sum"), g("gpu__time_duration.sum__unit"))
    print(f"| {rec['kernel']} | {t:.1f} | {rd:.1f} | {wr:.2f} | {float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
          f"{float(g('sm__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {float(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | "
          f"{float(g('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | {g('launch__registers_per_thread')} | "
          f"{g('launch__grid_size')} x {g('launch__block_size')} | {float(g('smsp__inst_executed.sum')) / 1e6:.1f} |")
    if "fit_" in rec["kernel"]:
        tot["fit_r"] += rd; tot["fit_w"] += wr
    for key in ("bin", "scatter"):
        if f"rpw_{key}_kernel" in rec["kernel"] and "compact" not in rec["kernel"]:
            per[key] = (rd + wr) * 1e6 / points
traffic = {"source": f"profiles/{raw.name} (ncu --set full, bench.py --scans 64, {points / 1e6:.2f} M points per launch group)",
           "fit_phase_dram_bytes_per_point": (tot["fit_r"] + tot["fit_w"]) * 1e6 / points,
           "fit_phase_dram_read_mb": tot["fit_r"], "fit_phase_dram_write_mb": tot["fit_w"], "points": points,
           "bin_dram_bytes_per_point": per.get("bin"), "scatter_dram_bytes_per_point": per.get("scatter")}
Path(raw.parent / "fit_traffic.json").write_text(json.dumps(traffic, indent=1))
print("\nfit phase DRAM bytes per input point: %.2f; bin %.2f; scatter %.2f" % (traffic["fit_phase_dram_bytes_per_point"], per.get("bin", 0), per.get("scatter", 0)))
