#!/bin/bash
# Two ncu passes over the same short bench command (run on the GPU box, one GPU):
#   1. launch list with per-launch device time   -> gpurun_out/$1_launches.csv
#   2. --set full (source + SASS) of ONE 32-frame batch step -> gpurun_out/$1_full.ncu-rep
# Numbers printed by a run under ncu are never bench values.
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --pool 32 --no-cpu-baseline --no-tracked --verify 0"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# first launch of the 5th 32-frame step (setup passes and warm-up come first); the launch ID is ncu's own count
SKIP=$(python - "$TAG" <<'P'
import csv, sys
lines = [l for l in open("gpurun_out/%s_launches.csv" % sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
ids = [int(r["ID"]) for r in rows if r["Kernel Name"].startswith("k_unpack") and r["Grid Size"].replace(" ", "").endswith(",64,1)")]
print(ids[4] if len(ids) > 4 else ids[-1])
P
)
echo "full capture from launch $SKIP" >> gpurun_out/${TAG}_ncu1.log
ncu --set full --clock-control none --import-source on --launch-skip $SKIP --launch-count 29 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
