#!/bin/bash
# A/B of two builds of libsvo_b200 on the GPU box in ONE gpurun call (one GPU):
#   1. the GPU parity tests on the in-tree build; if they fail, the extraction tests again with each of the new
#      kernels switched off (SVO_B200_RESIZE_QUADS / HARRIS8 / BLUR_MARGIN / BLUR_PACK = 0) to name the culprit
#   2. bench.py (resident-input value + host-chained e2e, last batches verified) on the in-tree build and on the build
#      named by $1 (default stereo-semantic-vo_b200/libsvo_b200_base.so), alternating, two runs each
#   3. the ncu launch list of the in-tree build (per-kernel device time)
# usage: tools/ab_builds.sh <tag> [other.so]
set -u
TAG=${1:-ab}
OTHER=${2:-$PWD/stereo-semantic-vo_b200/libsvo_b200_base.so}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
RC=$?
tail -5 gpurun_out/${TAG}_tests.log
if [ $RC -ne 0 ]; then
    for v in SVO_B200_RESIZE_QUADS SVO_B200_HARRIS8 SVO_B200_BLUR_MARGIN SVO_B200_BLUR_PACK; do
        echo "== $v=0" >> gpurun_out/${TAG}_tests.log
        env $v=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 >> gpurun_out/${TAG}_tests.log
    done
    echo "== all four off" >> gpurun_out/${TAG}_tests.log
    SVO_B200_RESIZE_QUADS=0 SVO_B200_HARRIS8=0 SVO_B200_BLUR_MARGIN=0 SVO_B200_BLUR_PACK=0 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 >> gpurun_out/${TAG}_tests.log
fi
BARGS="--steps 40 --warmup 4 --no-cpu-baseline --no-tracked"
for i in 1 2; do
    timeout 300 python bench.py $BARGS > gpurun_out/${TAG}_new_$i.json 2> gpurun_out/${TAG}_new_$i.err
    [ -f "$OTHER" ] && SVO_B200_LIB=$OTHER timeout 300 python bench.py $BARGS > gpurun_out/${TAG}_base_$i.json 2> gpurun_out/${TAG}_base_$i.err
done
python - "$TAG" <<'P'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_*_?.json" % sys.argv[1])):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f  e2e %.0f  p50 %.4f  verified %s  stages %s" % (l["value"], l["e2e"]["value"], l["p50_ms_per_frame_single"], l.get("verified_frames"),
              {k: round(v, 3) for k, v in l["stage_ms_per_step"].items()}))
    except Exception as e:
        print(f, "unreadable:", e)
P
CMD="python bench.py --steps 2 --warmup 3 --pool 32 --no-cpu-baseline --no-tracked --verify 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
python tools/ncu_step.py gpurun_out/${TAG}_launches.csv 32 > gpurun_out/${TAG}_step_summary.txt 2>&1
head -40 gpurun_out/${TAG}_step_summary.txt
