"""SASS opcode histogram per kernel of libmnemo_cuda.so's objects (profiles hygiene; also used by tests/test_sass.py).
usage: python tools/sass_histogram.py [--md profiles/<name>.md]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "mnemophonix_b200", "_obj")


def histograms():
    """{kernel name (demangled, short): Counter(opcode -> count)} over every .o under mnemophonix_b200/_obj."""
    out = {}
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith(".o"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
        cur = None
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = out.setdefault(name.split("(")[0].replace("mnemo::", "").replace("void ", ""), collections.Counter())
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
            if m and cur is not None:
                cur[m.group(1).split(".")[0]] += 1
    return out


if __name__ == "__main__":
    h = histograms()
    lines = ["# SASS opcode histogram per kernel (cuobjdump -sass of mnemophonix_b200/_obj/*.o, sm_100a)", ""]
    for k in sorted(h):
        c = h[k]
        total = sum(c.values())
        top = ", ".join(f"{op} {n}" for op, n in c.most_common(14))
        lines.append(f"* `{k}` — {total} instructions: {top}")
    text = "\n".join(lines) + "\n"
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write(text)
    print(text)
