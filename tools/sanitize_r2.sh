#!/bin/bash
# compute-sanitizer over the round-2 kernels (run on the GPU box): the device-resident tracker (track.cu), the compact
# result copies and the tensor-core Hamming tiles (tcham.cu) through their GPU tests.
OUT=gpurun_out/sanitizers_r2.txt
SEL="tracked_sequences or ballast or compact_and_left"
echo "# compute-sanitizer (B200) over tests/test_gpu_track.py -k '$SEL' and tests/test_gpu_tcham.py" > $OUT
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" >> $OUT
  timeout 1200 compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_track.py tests/test_gpu_tcham.py -m gpu -q -x -k "$SEL or tc" > gpurun_out/san2_$tool.log 2>&1
  grep -E "passed|failed" gpurun_out/san2_$tool.log | tail -1 >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san2_$tool.log | tail -1 >> $OUT
  grep -E "=========     at |========= .*(Invalid|hazard|Uninitialized|Barrier)" gpurun_out/san2_$tool.log | sort | uniq -c | sort -rn | head -8 >> $OUT
done
cat $OUT
