#!/bin/bash
# Everything a round wants from its FIRST GPU call, in one gpurun (one box acquisition):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round_check.sh rNN'
# 1. the GPU test-suite, 2. train.run end to end vs the oracle (both configs), 3. bench lines of the four
# single-GPU workloads + the scoring leg + the reference arm, 4. the ncu launch list of the default bench and one
# `--set full` capture of its dominant kernel (each only after the same command exited 0 without ncu).
# Outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
tail -3 $out/${tag}_gpu_tests.log
for cfg in autorec omni; do python scripts/train_check.py $cfg > $out/${tag}_train_check_$cfg.log 2>&1; tail -1 $out/${tag}_train_check_$cfg.log; done
python bench.py > $out/${tag}_bench_ml10m.json 2> $out/${tag}_bench_ml10m.err || exit 1
tail -c 600 $out/${tag}_bench_ml10m.json
for w in ml1m jester ml20m; do python bench.py --workload $w > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; done
( time timeout 900 python bench.py --workload netflix --steps 20 --no-cpu-baseline > $out/${tag}_bench_netflix.json ) 2> $out/${tag}_bench_netflix.err
python bench.py --mode score > $out/${tag}_bench_score.json 2> $out/${tag}_bench_score.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 4 --warmup 3 > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_row_update -s 6 -c 1 -f -o $out/${tag}_prof_rowupd \
    python bench.py --steps 4 --warmup 3 > $out/${tag}_ncu_rowupd.log 2>&1
ncu -i $out/${tag}_prof_rowupd.ncu-rep --page raw --csv > $out/${tag}_prof_rowupd_raw.csv 2>/dev/null
ls -la $out | tail -20
