#!/bin/bash
# Final captures of a round on the GPU box (one GPU): GPU tests, smoke, the two ncu passes, the bench lines of both arms,
# the other BASELINE configurations and the sanitizers.  usage: tools/final_capture.sh <tag>
set -u
TAG=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 bash tools/ncu_capture.sh ${TAG} > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
timeout 600 python tools/ncu_traffic.py gpurun_out/${TAG}_traffic.json gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_traffic.txt 2>&1; tail -3 gpurun_out/${TAG}_traffic.txt
timeout 600 python tools/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
python tools/ncu_step.py gpurun_out/${TAG}_launches.csv 32 > gpurun_out/${TAG}_step_summary.txt 2>&1
timeout 600 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench_line.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference_line.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-300 gpurun_out/${TAG}_bench_reference_line.json
timeout 600 python tools/other_configs.py gpurun_out/${TAG}_other_configs.json
timeout 900 bash tools/sanitize.sh > gpurun_out/${TAG}_sanitize.log 2>&1; cp gpurun_out/sanitizers.txt gpurun_out/${TAG}_sanitizers.txt; tail -20 gpurun_out/${TAG}_sanitizers.txt
