"""Instruction-count summary of the shipped library's SASS: per kernel, how many tcgen05 MMA
(UTCHMMA, .2CTA = cta_group::2), TMEM load / store (LDTM / STTM), TMA load (UTMALDG), MUFU and
vector-reduction instructions it contains -- the static proof that the hot kernels are
Blackwell-native (B200_PROFILING.md).

    python tools/sass_counts.py [robustvlm_b200/_lib/librvlm_b200.so] > profiles/r02_sass_counts.txt
"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2",
        "REDG", "STG.E.ENL2.256", "HMMA", "total"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "robustvlm_b200/_lib/librvlm_b200.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("rvlm::", "")
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = counts.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for k in KEYS:
                if k != "total" and (op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA.2CTA"))):
                    cur[k] += 1
    print(f"# {lib}: arch {sorted(arch)}; SASS instruction counts per kernel (static)")
    print(" ".join(f"{k:>9}" for k in ["UTCHMMA", "(.2CTA)", "LDTM", "STTM", "UTMALDG", "UTMASTG", "MUFU.EX2", "REDG", "HMMA", "total"]) + "  kernel")
    tot = collections.Counter()
    for name, c in counts.items():
        tot.update(c)
        if c["total"] < 50:
            continue
        print(" ".join(f"{c[k]:9d}" for k in ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "MUFU.EX2", "REDG", "HMMA", "total"]) + "  " + name[:120])
    print(" ".join(f"{tot[k]:9d}" for k in ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "MUFU.EX2", "REDG", "HMMA", "total"]) + "  TOTAL")


if __name__ == "__main__":
    main()
