"""Per-source-line hot spots of one kernel launch in an .ncu-rep: joins ncu's SASS page (instructions executed,
stall samples per instruction) with nvdisasm's line table of the cubin inside the shared library.

usage: python scripts/sass_hotspots.py REP KERNEL_REGEX [launch_index] [top_n]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "leica_point_cloud_processing_b200", "libgicp_b200.so")


def line_table(kernel_regex):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    tables = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
        fn, loc = None, ("?", 0)
        for ln in out.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                fn, loc = m.group(1), ("?", 0)
                tables[fn] = {}
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m and fn:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
            if m and fn:
                tables[fn][int(m.group(1), 16)] = loc
    return tables


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                         capture_output=True, text=True).stdout
    allrows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
    rows = allrows[starts[2 * launch]:starts[2 * launch + 1]]  # ncu prints every launch twice (two views)
    kname = rows[0][1]
    hdr = rows[1]
    ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = int(rows[2][ia], 16)
    tables = line_table(rx)
    # pick the table whose demangled-ish name matches: same instruction count
    ninstr = len(rows) - 2
    cands = [(fn, t) for fn, t in tables.items() if re.search(rx, fn) and abs(len(t) - ninstr) <= 2]
    if not cands:
        cands = [(fn, t) for fn, t in tables.items() if re.search(rx, fn)]
    # disambiguate template instances by launch order heuristics: take the one with equal length first
    want = re.sub(r"[^A-Za-z0-9]", "", kname.split("(")[0].split("::")[-1])  # e.g. correspondence_kerneldoublebool0
    fn, table = cands[0]
    for f, t in cands:
        tag = ("Id" if "<double" in kname else "If" if "<float" in kname else "") + ("Lb1" if "(bool)1" in kname else "Lb0" if "(bool)0" in kname else "")
        if tag and tag in f:
            fn, table = f, t
            break
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        loc = table.get(off, ("?", 0))
        v = (int(r[ie]), int(r[it]), int(r[isamp]))
        for k in range(3):
            agg[loc][k] += v[k]
            tot[k] += v[k]
    print(f"# {kname[:100]}\n# table {fn[:80]}  warp-instr {tot[0]:,}  thread-instr {tot[1]:,}  avg lanes {tot[1] / max(tot[0], 1):.1f}  samples {tot[2]}")
    print(f"{'file:line':28s} {'warp-instr%':>11s} {'lanes':>6s} {'samples%':>9s}")
    for loc, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        src = ""
        path = os.path.join(ROOT, "leica_point_cloud_processing_b200", "csrc", loc[0])
        if os.path.exists(path) and loc[1] > 0:
            src = open(path).read().splitlines()[loc[1] - 1].strip()[:90]
        print(f"{loc[0] + ':' + str(loc[1]):28s} {100 * v[0] / tot[0]:10.1f}% {v[1] / max(v[0], 1):6.1f} {100 * v[2] / max(tot[2], 1):8.1f}%  {src}")


if __name__ == "__main__":
    main()
