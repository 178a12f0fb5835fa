#!/bin/bash
# compute-sanitizer (memcheck, initcheck, racecheck, synccheck) over the GPU parity tests that reach every kernel.
# usage: scripts/gpu_sanitize.sh <out-file>
OUT=${1:-gpurun_out/sanitizer.txt}
SEL='golden or text_batch or empty or odd or pairing or random_programs or synthetic or homopolymer'
echo "# compute-sanitizer, 1x B200: python -m pytest tests/test_gpu_parity.py -m gpu -k '$SEL' under each tool" > $OUT
for tool in memcheck initcheck racecheck synccheck; do
  echo "== $tool" >> $OUT
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" 2>&1 \
    | grep -E "passed|failed|error|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|at 0x|by thread" | head -40 >> $OUT
done
cat $OUT
