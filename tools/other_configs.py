#!/usr/bin/env python3
"""Run the non-headline BASELINE.json configurations (and the opt-in octree mode) through bench.py and collect
profiles/r1_other_configs.json-style entries.  usage (GPU box): python tools/other_configs.py out.json"""
import json
import subprocess
import sys

RUNS = {
    "configs[2] K/4000": "python bench.py --no-cpu-baseline --features 4000",
    "configs[3] H/8000": "python bench.py --no-cpu-baseline --width 2560 --height 720 --features 8000 --batch 8 --pool 24 --steps 30",
    "configs[1] with the opt-in octree distribution": "python bench.py --no-cpu-baseline --distribution octree",
}
out = {}
for name, cmd in RUNS.items():
    txt = subprocess.run(cmd.split(), capture_output=True, text=True).stdout
    line = [l for l in txt.splitlines() if l.startswith("{")]
    if not line:
        out[name] = {"command": cmd, "error": "no JSON line"}
        continue
    d = json.loads(line[0])
    out[name] = {"command": cmd, "value_fps": d["value"], "e2e_fps": d["e2e"]["value"],
                 "p50_ms_per_frame_single": d.get("p50_ms_per_frame_single"), "frames_per_step": d["config"]["frames_per_step"],
                 "stage_ms_per_step": d.get("stage_ms_per_step"), "kernel_ms_per_launch": d.get("kernel_ms_per_launch"),
                 "clocks": d.get("clocks")}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps({k: (v.get("value_fps"), v.get("e2e_fps"), v.get("p50_ms_per_frame_single")) for k, v in out.items()}))
