#!/bin/bash
# One GPU-box visit: parity tests, bench (default + A/B variants), ncu launch list + full capture of the chain.
# Everything lands in gpurun_out/<tag>/.   usage: scripts/gpu_round.sh <tag> [tests|notests]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest.log
  tail -5 $OUT/pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; head -c 600 $OUT/bench.json; echo
VARIANTS="${VARIANTS:---emit g16;--homo v1;--no-numa}"
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  name=$(echo $v | tr -d ' -')
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-files $v > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $v exit $?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-files --no-e2e --batches 2 --batch-pairs 1000000 > $OUT/ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_emit_stage|k_align|k_scan}" -c ${NCU_COUNT:-8} -o $OUT/new_kernels \
  python bench.py --steps 1 --warmup 0 --no-cpu --no-files --no-e2e --batches 1 --batch-pairs 1000000 > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"
# gpurun brings back at most 64 MiB: keep the per-launch summary, drop a report that would not fit
python scripts/ncu_summary.py $OUT/new_kernels.ncu-rep $OUT/ncu_summary.csv > /dev/null 2>&1
if [ "$(stat -c %s $OUT/new_kernels.ncu-rep 2>/dev/null || echo 0)" -gt 40000000 ]; then rm -f $OUT/new_kernels.ncu-rep; fi
ls -la $OUT
