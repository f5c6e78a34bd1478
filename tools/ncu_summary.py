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


def table(path, hbm_peak_gbs=6532.2):
    """one line per distinct kernel (mean over its captured launches): time, DRAM bytes, achieved
    GB/s and its fraction of the measured HBM peak (MEASURED_PEAKS.json hbm_gbs), tensor-pipe
    utilisation, issue-slot utilisation"""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(d, k):
        try:
            v = float(d[k].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")
        u = units[hdr.index(k)]
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3,
                    "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    agg = collections.OrderedDict()
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("rvlm::", "")
        name = re.sub(r"<unnamed>::", "", name)
        key = (name, d.get("Grid Size", ""))
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += val(d, "gpu__time_duration.sum")
        a[2] += val(d, "dram__bytes_read.sum")
        a[3] += val(d, "dram__bytes_write.sum")
        a[4] += val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        a[5] += val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    print(f"# {path}: per-kernel means; HBM peak = {hbm_peak_gbs} GB/s (MEASURED_PEAKS.json); ncu "
          "--clock-control none, kernels serialised (cold L2)")
    print(f"{'n':>3} {'time_us':>9} {'rd_MB':>8} {'wr_MB':>8} {'GB/s':>7} {'of_HBM':>6} {'tensor%':>7} "
          f"{'issue%':>6}  kernel  grid")
    for (name, grid), (n, t, rd, wr, tp, ia) in agg.items():
        t, rd, wr, tp, ia = t / n, rd / n, wr / n, tp / n, ia / n
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else float("nan")
        print(f"{n:3d} {t:9.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {gbs:7.0f} {gbs / hbm_peak_gbs:6.2f} "
              f"{tp:7.1f} {ia:6.1f}  {name[:110]}  {grid}")


if __name__ == "__main__":
    {"launches": launches, "report": report, "table": table}[sys.argv[1]](sys.argv[2])
