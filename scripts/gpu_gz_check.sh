#!/bin/bash
# gzip codecs on the GPU box: tests, kernel time of the inflate (ncu, one mate of 1 M reads), whole-file .gz run.  usage: scripts/gpu_gz_check.sh <tag>
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_gz.py -x -q > $OUT/pytest.log 2>&1; echo pytest exit $?; tail -3 $OUT/pytest.log
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_gz_inflate" -c 2 --csv --log-file $OUT/gz_launches.csv python scripts/gz_one_batch.py --pairs 1000000 --reps 1 > $OUT/run.log 2>&1
grep -o "\"[a-z_.]*\",\"[a-zA-Z%/]*\",\"[0-9.,]*\"$" $OUT/gz_launches.csv | head -4
timeout 400 python scripts/bench_files.py --pairs ${2:-8000000} --gpus 1 --threads 16 --variants gz --out $OUT/files.json > $OUT/files.log 2>&1; grep "pairs_per_s\|wall_s" $OUT/files.json | head -2
