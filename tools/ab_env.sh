#!/bin/bash
# A/B of RUN-TIME switches of the library on the bench workload (through gpurun):
#   bash tools/ab_env.sh <tag> name1:ENV=VAL[,ENV2=VAL] name2:...     ("name:" alone = defaults)
# One short bench.py run per spec on the same box; prints img/s and the in-step breakdown side by side.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; shift
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  envs=${envs//,/ }
  env $envs timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-modes \
    --no-torch-baseline --no-train-step > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err \
    || { echo "$name FAILED"; tail -5 gpurun_out/${TAG}_${name}.err; }
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
print("%-34s" % "variant" + "".join("%10s" % n for n in names))
print("%-34s" % "img/s" + "".join("%10.1f" % rows[n]["value"] for n in names))
print("%-34s" % "e2e img/s" + "".join("%10.1f" % rows[n]["e2e"]["value"] for n in names))
print("%-34s" % "sm_mhz" + "".join("%10.0f" % rows[n]["clocks"]["sm_mhz"] for n in names))
keys = []
for n in names:
    for k in rows[n].get("kernel_breakdown", {}):
        if k not in keys:
            keys.append(k)
for k in keys:
    print("%-34s" % k[:34] + "".join(
        "%10.1f" % rows[n].get("kernel_breakdown", {}).get(k, {}).get("ms_per_step", float("nan"))
        for n in names))
PY
