"""Attribute an ncu report's warp-stall samples to SOURCE LINES of one kernel.

  python tools/ncu_regions.py <report.ncu-rep> <object.o> <kernel-substring> [line ranges...]

The SASS page of the report gives samples per instruction (in program order); `nvdisasm
--print-line-info` on the same object's cubin gives the source line of each instruction of the
same function, in the same order.  Joining them by instruction index yields samples per source
line, summed over the line ranges given as name:first-last (inclusive).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
    return list(csv.DictReader(io.StringIO("\n".join(out[start:]))))


def line_table(obj, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    # -gi prints the inlining chain innermost first; the LAST annotation of a block is the
    # outermost call site, i.e. the line of the kernel body the instruction belongs to
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(d, cub)],
                         capture_output=True, text=True).stdout.splitlines()
    lines, cur, inside = [], 0, False
    for l in txt:
        if l.startswith(".text."):
            inside = kernel in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            f, n = (m.group(3), m.group(4)) if m.group(3) else (m.group(1), m.group(2))
            cur = int(n) if f.endswith(".cu") else -int(n)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    return lines


def main():
    rep, obj, kernel = sys.argv[1:4]
    rows = sass_rows(rep)
    lt = line_table(obj, kernel)
    print(f"{len(rows)} SASS rows in report, {len(lt)} instructions in {kernel}")
    per = {}
    for i, r in enumerate(rows):
        ln = lt[i] if i < len(lt) else 0
        per[ln] = per.get(ln, 0) + int(r["# Samples"] or 0)
    tot = sum(per.values())
    for spec in sys.argv[4:]:
        name, rg = spec.split(":")
        a, b = map(int, rg.split("-"))
        s = sum(v for k, v in per.items() if a <= k <= b)
        print(f"{name:40s} {s:7d} {100.0 * s / tot:5.1f}%")
    print("top lines:")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:30]:
        print(f"  line {k:5d}: {v:6d} {100.0 * v / tot:5.1f}%")


if __name__ == "__main__":
    main()
