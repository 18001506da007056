"""Turns an ncu report of one step of the path (tools/gpu_prof_step.py under `ncu --set full --import-source on`) into
the tracked summaries:

    python profiles/summarize.py gpurun_out/r02_step512.ncu-rep r02 512

  profiles/<tag>_path_summary.md        per launch: time, DRAM bytes, issue / warps active, registers, grid, the warp-state
                                        breakdown (stall reasons per issued instruction), shared-memory bank conflicts,
                                        local-memory (spill) instructions
  profiles/<tag>_path_full_summary.json the same numbers, machine readable
  profiles/<tag>_fit_hotspots.md        per fit kernel: opcode mix and the 20 SASS lines with the most stall samples
  profiles/fit_traffic.json             DRAM bytes per input point of the fit phase, bin and scatter (bench.py's
                                        roofline.traffic), from THIS capture's batch size
"""
import collections, csv, io, json, re, subprocess, sys
from pathlib import Path

rep = Path(sys.argv[1]); tag = sys.argv[2]
scans = int(sys.argv[3]) if len(sys.argv) > 3 else 512
points = scans * 120000
out_dir = Path(__file__).resolve().parent


def ncu(*args):
    return subprocess.run(["ncu", "-i", str(rep), *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, data = raw[0], raw[1], raw[2:]
col = {h: i for i, h in enumerate(hdr)}
STALLS = ["no_instruction", "barrier", "short_scoreboard", "long_scoreboard", "math_pipe_throttle", "wait", "branch_resolving", "mio_throttle",
          "lg_throttle", "dispatch_stall", "not_selected", "imc_miss", "membar", "sleeping"]
keep = {"time_us": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
        "block": "launch__block_size", "warp_inst": "smsp__inst_executed.sum", "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "local_ld_inst": "smsp__inst_executed_op_local_ld.sum", "local_st_inst": "smsp__inst_executed_op_local_st.sum",
        "occ_limit_smem": "launch__occupancy_limit_shared_mem", "occ_limit_regs": "launch__occupancy_limit_registers"}
for s in STALLS:
    keep["stall_" + s] = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"


def num(row, metric):
    if metric not in col:
        return None
    v, u = row[col[metric]], units[col[metric]]
    try:
        v = float(v.replace(",", ""))
    except ValueError:
        return None
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v * scale


recs = []
for r in data:
    name = re.sub(r"^void |rpw::|\(.*$", "", r[col["Kernel Name"]])
    rec = {"kernel": name}
    for k, m in keep.items():
        rec[k] = num(r, m)
    recs.append(rec)
(out_dir / f"{tag}_path_full_summary.json").write_text(json.dumps(recs, indent=1))

lines = [f"# {tag}: one step of the path under `ncu --set full` ({scans} C2 scans = {points / 1e6:.1f} M points resident; every buffer exceeds the 126 MB L2)", "",
         "Each launch is replayed alone (the seven fit classes run one after the other here, concurrently in production), caches flushed between replays: "
         "compare shares and per-kernel ratios, not absolute times.", "",
         "| kernel | grid x block | regs | time us | DRAM R MB | DRAM W MB | DRAM % | warps active % | issue active % | warp instr M | smem bank conflicts k | local ld/st instr k |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
f = lambda v, d=1: "-" if v is None else f"{v:.{d}f}"
for rec in recs:
    lines.append(f"| {rec['kernel']} | {f(rec['grid'], 0)} x {f(rec['block'], 0)} | {f(rec['regs'], 0)} | {f(rec['time_us'])} | {f((rec['dram_read'] or 0) / 1e6)} | "
                 f"{f((rec['dram_write'] or 0) / 1e6, 2)} | {f(rec['dram_pct'])} | {f(rec['warps_active_pct'])} | {f(rec['issue_active_pct'])} | {f((rec['warp_inst'] or 0) / 1e6)} | "
                 f"{f((rec['smem_bank_conflicts'] or 0) / 1e3)} | {f((rec['local_ld_inst'] or 0) / 1e3)} / {f((rec['local_st_inst'] or 0) / 1e3)} |")
lines += ["", "Warp states: average number of warps per scheduler in each stall reason per issued instruction (`smsp__average_warps_issue_stalled_*_per_issue_active`).", "",
          "| kernel | " + " | ".join(STALLS) + " |", "|---|" + "---|" * len(STALLS)]
for rec in recs:
    lines.append(f"| {rec['kernel']} | " + " | ".join(f(rec["stall_" + s], 2) for s in STALLS) + " |")

fit = [r for r in recs if "fit_" in r["kernel"]]
fit_r = sum(r["dram_read"] or 0 for r in fit); fit_w = sum(r["dram_write"] or 0 for r in fit)
per = {}
for key in ("bin", "scatter"):
    for r in recs:
        if f"rpw_{key}_kernel" in r["kernel"] and "compact" not in r["kernel"]:
            per[key] = ((r["dram_read"] or 0) + (r["dram_write"] or 0)) / points
tot_inst = sum(r["warp_inst"] or 0 for r in fit)
issue_floor_ms = tot_inst / (148 * 4 * 1.965e9) * 1e3
lines += ["", f"Fit phase: {tot_inst / 1e6:.0f} M warp instructions per step = {tot_inst * 32 / points:.0f} thread instructions per input point; at 148 SMs x 4 "
          f"schedulers x 1.965 GHz that is {issue_floor_ms:.3f} ms of pure issue time.  DRAM traffic of the fit phase {(fit_r + fit_w) / points:.2f} B per input point "
          f"(read {fit_r / 1e6:.0f} MB, write {fit_w / 1e6:.1f} MB), bin {per.get('bin', 0):.2f}, scatter {per.get('scatter', 0):.2f}."]
(out_dir / f"{tag}_path_summary.md").write_text("\n".join(lines) + "\n")
traffic = {"source": f"profiles/{tag}_path_full_summary.json (ncu --set full, tools/gpu_prof_step.py {scans}: {points / 1e6:.2f} M points per launch group)",
           "fit_phase_dram_bytes_per_point": (fit_r + fit_w) / points, "fit_phase_dram_read_mb": fit_r / 1e6, "fit_phase_dram_write_mb": fit_w / 1e6, "points": points,
           "bin_dram_bytes_per_point": per.get("bin"), "scatter_dram_bytes_per_point": per.get("scatter"),
           "fit_phase_warp_instructions": tot_inst, "fit_phase_issue_floor_ms": issue_floor_ms}
(out_dir / "fit_traffic.json").write_text(json.dumps(traffic, indent=1))

# ---- per fit kernel: opcode mix and the hottest SASS lines (stall samples) ------------------------------------------
hot = [f"# {tag}: where the fit kernels' warps wait (ncu source page, SASS level, {scans} scans)", "",
       "Per kernel: share of executed warp instructions by opcode, and the 20 SASS instructions with the most warp-stall samples "
       "(`Warp Stall Sampling (All Samples)`), with the instructions before them for context."]
n_fit = sum(1 for r in recs if "rpw_fit_roots" in r["kernel"])
for k in range(n_fit):
    txt = ncu("--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:rpw_fit_roots", "--launch-skip", str(k), "--launch-count", "1")
    rows = list(csv.reader(io.StringIO(txt)))
    try:
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    except StopIteration:
        continue
    name = next((r[1] for r in rows[:hi] if r and r[0] == "Kernel Name"), "rpw_fit_roots_kernel")
    h = {c: i for i, c in enumerate(rows[hi])}
    body = rows[hi + 1:]
    ops, stall_by_op, tot, ts = collections.Counter(), collections.Counter(), 0, 0
    parsed, seen = [], set()
    for r in body:
        if r and r[0] in seen:  # (the page lists every address twice)
            continue
        seen.add(r[0] if r else None)
        try:
            n = int(r[h["Instructions Executed"]]); s = int(r[h["Warp Stall Sampling (All Samples)"]])
        except (ValueError, IndexError):
            continue
        src = r[h["Source"]].strip()
        tok = src.split()
        op = (tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")).split(".")[0]
        ops[op] += n; stall_by_op[op] += s; tot += n; ts += s
        parsed.append((s, n, src))
    fitrec = [r for r in recs if "rpw_fit_roots" in r["kernel"]][k]
    hot += ["", f"## launch {k}: {name} <{f(fitrec['block'], 0)} threads>, grid {f(fitrec['grid'], 0)}, {f(fitrec['time_us'])} us alone, issue active {f(fitrec['issue_active_pct'])} %", "",
            "opcode mix (share of warp instructions / share of stall samples): " +
            ", ".join(f"{op} {100 * n / max(1, tot):.1f}/{100 * stall_by_op[op] / max(1, ts):.1f}" for op, n in ops.most_common(14)), "",
            "| stall samples % | executed (k warp instr) | SASS |", "|---|---|---|"]
    for s, n, src in sorted(parsed, reverse=True)[:20]:
        hot.append(f"| {100 * s / max(1, ts):.2f} | {n / 1e3:.0f} | `{src[:110]}` |")
(out_dir / f"{tag}_fit_hotspots.md").write_text("\n".join(hot) + "\n")
print("\n".join(lines[-3:]))
