#!/usr/bin/env python
"""Research tool: where a kernel's executed warp instructions are, by source line.

  python tools/sass_hotspots.py <report.ncu-rep> <kernel regex> <launch skip> [top N]

Joins the per-instruction counters of an `ncu --set full --import-source on` capture (`--page source --csv`) with the
line table of the same kernel in nrenderer_b200/libnrcuda.so (`nvdisasm -gi`), instruction by instruction, and prints the
source lines (innermost frame inside this repository) that execute the most warp instructions / collect the most stall samples.
"""
import collections, csv, io, os, re, subprocess, sys, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def ncu_sass(rep, regex, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex, "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    rep_at = next((i for i in range(1, len(data)) if data[i] and data[i][0] == data[0][0]), None)   # the CSV lists the kernel twice
    if rep_at: data = data[:rep_at]
    return name, hdr, data

def line_table(mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "nrenderer_b200", os.environ.get("NRCU_SO", "libnrcuda.so"))], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    start = [i for i, l in enumerate(txt) if l.startswith(".text.") and mangled_hint in l][0]
    table, frames = [], []
    for l in txt[start + 1:]:
        if l.startswith(".text.") or l.lstrip().startswith(".section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            frames.append((m.group(1), int(m.group(2)))); continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            if frames: cur = frames
            frames = []
            table.append((l.split("*/", 1)[1].strip(), cur))
    return table

def main():
    rep, regex, skip = sys.argv[1], sys.argv[2], int(sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    name, hdr, data = ncu_sass(rep, regex, skip)
    hint = os.environ.get("NRCU_MANGLED", regex)
    table = line_table(hint)
    iI, iN, iS = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    iL = hdr.index("stall_long_sb")
    if len(table) != len(data):
        print(f"warning: {len(table)} instructions in the library, {len(data)} in the report (different build?)")
        if abs(len(table) - len(data)) > 2: sys.exit(1)
    src_cache = {}
    def text(f, ln):
        if f not in src_cache:
            try: src_cache[f] = open(f).read().splitlines()
            except OSError: src_cache[f] = []
        L = src_cache[f]
        return L[ln - 1].strip() if 0 < ln <= len(L) else ""
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    total = [0, 0]
    for (sass, frames), r in zip(table, data):
        ex, smp, lsb = int(r[iI]), int(r[iN]), int(r[iL])
        inner = next((fr for fr in frames if fr[0].startswith(REPO)), frames[-1])
        outer = frames[-1]
        a = agg[(inner, outer[1])]
        a[0] += ex; a[1] += smp; a[2] += lsb; a[3] += 1
        total[0] += ex; total[1] += smp
    print(name[:100]); print(f"executed warp instructions {total[0]/1e6:.1f} M, samples {total[1]}")
    byinner = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (inner, outer), a in agg.items():
        b = byinner[inner]
        for k in range(4): b[k] += a[k]
    print(f"{'% inst':>7} {'% smp':>6} {'longsb':>7} {'sass':>5}  source line")
    for inner, a in sorted(byinner.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{a[0]/total[0]*100:7.2f} {a[1]/max(total[1],1)*100:6.2f} {a[2]:7d} {a[3]:5d}  {os.path.basename(inner[0])}:{inner[1]}  {text(*inner)[:110]}")

if __name__ == "__main__":
    main()
