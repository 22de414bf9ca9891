"""Per-source-line hot spots of one kernel from an .ncu-rep (source page, cuda,sass view).
usage: python tools/ncu_hotspots.py report.ncu-rep [top_n] [kernel-name regex] > profiles/<name>_source_hotspots.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, *flt, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
items, fpath, hdr = [], None, None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name") or hdr is None or not r[0]:
        continue                                   # rows without a line number are the SASS lines under a source line
    try:
        i_inst, i_smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        items.append((float(r[i_inst]), float(r[i_smp]), fpath, r[0], r[1].strip()[:120]))
    except (ValueError, IndexError):
        pass
tot = sum(i[0] for i in items) or 1
ts = sum(i[1] for i in items) or 1
print(f"total warp-inst {tot:.0f} samples {ts:.0f}")
for n, sm, f, l, src in sorted(items, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}% inst {100 * sm / ts:5.1f}% smp  {f}:{l:>4}  {src}")
