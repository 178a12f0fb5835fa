#!/bin/bash
# ncu evidence for profiles/: launch list of the bench step, --set full of one whole chain (1 M pairs) and of the gzip
# kernels (one BGZF -> gzip batch of 1 M pairs).   usage: scripts/gpu_profile.sh <tag>
set -u
OUT=gpurun_out/$1; mkdir -p $OUT
python -c "import bench; print(bench.kernel_sources_hash())" > $OUT/sources_hash.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-files --no-e2e --batches 2 --batch-pairs 1000000 > $OUT/ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -c 40 -o $OUT/chain \
  python bench.py --steps 1 --warmup 0 --no-cpu --no-files --no-e2e --batches 1 --batch-pairs 1000000 > $OUT/ncu_chain.log 2>&1
echo "ncu chain exit $?"
python scripts/ncu_summary.py $OUT/chain.ncu-rep $OUT/chain_summary.csv > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gz" -c 14 -o $OUT/gz \
  python scripts/gz_one_batch.py --pairs 1000000 --reps 1 > $OUT/ncu_gz.log 2>&1
echo "ncu gz exit $?"
python scripts/ncu_summary.py $OUT/gz.ncu-rep $OUT/gz_summary.csv > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/gz_launches.csv python scripts/gz_one_batch.py --pairs 1000000 --reps 1 > /dev/null 2>&1
for f in chain gz; do if [ "$(stat -c %s $OUT/$f.ncu-rep 2>/dev/null || echo 0)" -gt 30000000 ]; then rm -f $OUT/$f.ncu-rep; fi; done
ls -la $OUT
