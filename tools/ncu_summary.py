"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv  > profiles/r01_launches_summary.txt
  python tools/ncu_summary.py report   gpurun_out/r01_gemm.ncu-rep  > profiles/r01_gemm_ncu.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size",
    "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def short(name):
    m = re.search(r"gemm16_kernel<(\d+), *\(?bool\)?(\d), *\(?bool\)?(\d), *(?:rvlm::)?(\w+)>", name)
    if m:
        return f"gemm16<BN={m.group(1)},{m.group(4)}>"
    m = re.search(r"attn_tc_kernel<(\d)>", name)
    if m:
        return f"attn_tc<{['FWD', 'DQ', 'DKV'][int(m.group(1))]}>"
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("rvlm::", "")


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time "
          "(ncu: serialised, cold-cache; compare SHARES)")
    print(f"{'share':>7} {'total_us':>11} {'n':>5} {'avg_us':>9}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / tot * 100:6.2f}% {t:11.1f} {n:5d} {t / n:9.1f}  {k}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        print(f"== {short(d['Kernel Name'])}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"   {k} = {d[k]} {units[hdr.index(k)]}")
        stalls = [(h, float(d[h])) for h in hdr if "issue_stalled" in h and h.endswith(".ratio")
                  and "not_issued" not in h and d[h] not in ("", "n/a")]
        top = sorted(stalls, key=lambda kv: -kv[1])[:5]
        print("   top stalls: " + ", ".join(
            f"{h.split('issue_stalled_')[1].split('_per')[0]}={v:.2f}" for h, v in top))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
