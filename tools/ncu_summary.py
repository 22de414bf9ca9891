"""Summarise an .ncu-rep (raw page) into a markdown table + traffic.json entries.
usage: python tools/ncu_summary.py report.ncu-rep out.md [traffic.json] [--title "workload and build the capture is of"]"""
import csv, io, json, subprocess, sys
title = None
if "--title" in sys.argv:
    i = sys.argv.index("--title")
    title = sys.argv[i + 1]
    del sys.argv[i:i + 2]
rep, out_md = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "warp insts"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math")]
idx = {c: hdr.index(c) for c, _ in cols if c in hdr}
ki = hdr.index("Kernel Name")
lines = ["| kernel | " + " | ".join(n for c, n in cols if c in idx) + " |", "|---|" + "---|" * len(idx)]
traffic = {}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("mnemo::", "")
    vals = []
    for c, n in cols:
        if c not in idx: continue
        v, u = r[idx[c]], units[idx[c]]
        try: f = float(v.replace(",", "")); v = f"{f:.4g}"
        except ValueError: pass
        vals.append(f"{v} {u}".strip())
    lines.append(f"| {name} | " + " | ".join(vals) + " |")
    try:
        rd = float(r[hdr.index('dram__bytes_read.sum')].replace(",", "")) * scale.get(units[hdr.index('dram__bytes_read.sum')], 1)
        wr = float(r[hdr.index('dram__bytes_write.sum')].replace(",", "")) * scale.get(units[hdr.index('dram__bytes_write.sum')], 1)
        key = {"k0_resample_kernel": "k0_resample", "k1_spectral_kernel": "k1_spectral", "k2_fingerprint_kernel<1>": "k2_fingerprint",
               "k2_fingerprint_kernel<0>": "k2_fingerprint"}.get(name.strip(), name.strip())
        traffic[key] = rd + wr
    except Exception:
        pass
if title:
    lines = [title, ""] + lines
open(out_md, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3:
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
print("\n".join(lines))
