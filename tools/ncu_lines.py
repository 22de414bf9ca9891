"""Dev tool: aggregate an ncu report's per-SASS-instruction counts by CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: [0, 0, ""])
fname = ""
hdr = None
cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); sa = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie: continue
    if r[0]:      # a CUDA source line row (its own totals)
        cur = (fname, r[0]); agg[cur][2] = r[1].strip()[:100]
        try:
            agg[cur][0] += int(r[ie] or 0); agg[cur][1] += int(r[sa] or 0)
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total warp-inst", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/max(tot,1):5.1f}% inst {100*v[1]/max(tots,1):5.1f}% smp  {k[0]}:{k[1]:>4s}  {v[2]}")
