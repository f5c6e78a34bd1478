"""Top stall sites of one kernel from an ncu report's SASS source page.

  python tools/ncu_hot.py gpurun_out/x.ncu-rep [N]   -> the N SASS instructions with the most
  warp-stall samples, with their dominant stall reasons, plus totals per stall reason.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    print(lines[0][:200])
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    tot = {c: 0 for c in stall_cols}
    total = 0
    for i, r in enumerate(rows):
        r["_i"] = i
        r["_s"] = int(r["# Samples"] or 0)
        total += r["_s"]
        for c in stall_cols:
            tot[c] += int(r[c] or 0)
    print("total samples", total)
    print("by reason:", ", ".join(f"{c[6:]}={v}" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    for r in sorted(rows, key=lambda r: -r["_s"])[:n]:
        why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{r['_s']:6d} {100.0 * r['_s'] / total:5.1f}%  #{r['_i']:4d} {r['Source'][:70]:70s} "
              + " ".join(f"{k}={v}" for v, k in why if v))


if __name__ == "__main__":
    main()
