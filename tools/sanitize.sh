#!/bin/bash
# compute-sanitizer over the GPU parity tests that exercise every kernel family (run on the GPU box).
SEL="batch_pipeline or stages or greedy_vs or stereo or octree_extract or projection_windows or any_residency"
SEL2="batch_pipeline or greedy_vs or octree_extract_vs_oracle"
OUT=gpurun_out/sanitizers.txt
echo "# compute-sanitizer runs (B200) over tests/test_gpu_parity.py -k '$SEL' (racecheck/synccheck/initcheck: -k '$SEL2')" > $OUT
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" >> $OUT
  if [ $tool = memcheck ]; then K="$SEL"; else K="$SEL2"; fi
  timeout 900 compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/san_$tool.log 2>&1
  grep -E "passed|failed" gpurun_out/san_$tool.log | tail -1 >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san_$tool.log | tail -1 >> $OUT
  grep -E "=========     at |========= .*(Invalid|hazard|Uninitialized|Barrier)" gpurun_out/san_$tool.log | sort | uniq -c | sort -rn | head -8 >> $OUT
done
cat $OUT
