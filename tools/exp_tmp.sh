#!/bin/bash
set -u
mkdir -p gpurun_out
B="--steps 40 --warmup 4 --no-cpu-baseline --no-tracked"
P=$PWD/stereo-semantic-vo_b200
run() { tag=$1; shift; timeout 300 env "$@" > gpurun_out/e4_$tag.json 2> gpurun_out/e4_$tag.err; }
run cur_1 python bench.py $B
run w4_1 SVO_B200_LIB=$P/libsvo_b200_w4.so python bench.py $B
run cur_2 python bench.py $B
run w4_2 SVO_B200_LIB=$P/libsvo_b200_w4.so python bench.py $B
run band8 SVO_B200_FAST_BAND=8 python bench.py $B
run lanes5 python bench.py $B --lanes 5
run b40 python bench.py $B --batch 40
run b24l5 python bench.py $B --batch 24 --lanes 5
python - <<'P'
import glob, json
for f in sorted(glob.glob("gpurun_out/e4_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f  e2e %.0f p50 %.4f  verified %s  stages %s" % (l["value"], l["e2e"]["value"], l["p50_ms_per_frame_single"], l.get("verified_frames"),
              {k: round(v, 3) for k, v in l["stage_ms_per_step"].items()}))
    except Exception as e:
        print(f, "unreadable:", e)
P
