#!/bin/bash
# A/B of library builds on the bench workload (run through gpurun): for every
# robustvlm_b200/_lib/var_*/librvlm_b200.so (csrc/Makefile OUTDIR=... EXTRA=-D...) and the main
# build, a short bench.py run; prints img/s and the in-step per-kernel breakdown side by side.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-ab}
run() {  # name, lib, extra env
  local name=$1 lib=$2; shift 2
  env RVLM_LIB=$lib "$@" timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-modes \
    --no-torch-baseline --no-train-step > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err \
    || { echo "$name FAILED"; tail -5 gpurun_out/${TAG}_${name}.err; }
}
run main "$PWD/robustvlm_b200/_lib/librvlm_b200.so"
for d in robustvlm_b200/_lib/var_*/; do
  n=$(basename $d); n=${n#var_}
  if [ "$n" = base ]; then run $n "$PWD/$d/librvlm_b200.so" RVLM_ATTN_FWD_ONE=0
  else run $n "$PWD/$d/librvlm_b200.so"; fi
done
python - "$TAG" <<'PY'
import glob, json, os, sys
tag = sys.argv[1]
rows = {}
for f in sorted(glob.glob(f"gpurun_out/{tag}_*.json")):
    name = os.path.basename(f)[len(tag) + 1:-5]
    try:
        rows[name] = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(name, "unreadable", e)
names = list(rows)
print("%-28s" % "variant" + "".join("%10s" % n for n in names))
print("%-28s" % "img/s" + "".join("%10.1f" % rows[n]["value"] for n in names))
print("%-28s" % "sm_mhz" + "".join("%10.0f" % rows[n]["clocks"]["sm_mhz"] for n in names))
keys = []
for n in names:
    for k in rows[n].get("kernel_breakdown", {}):
        if k not in keys:
            keys.append(k)
for k in keys:
    print("%-28s" % k[:28] + "".join(
        "%10.1f" % rows[n].get("kernel_breakdown", {}).get(k, {}).get("ms_per_step", float("nan"))
        for n in names))
PY
