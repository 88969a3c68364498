#!/bin/bash
# round 2, call m: whole GPU suite, smoke, the driver's own default bench command (with other_workloads) + reference arm, launch list
out=gpurun_out; tag=${1:-r02m}
mkdir -p $out
timeout 2400 python -m pytest tests -q -m gpu --durations=15 > $out/${tag}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
tail -8 $out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
( time python bench.py > $out/${tag}_bench_default.json ) 2> $out/${tag}_bench_default.err
tail -c 400 $out/${tag}_bench_default.json; tail -4 $out/${tag}_bench_default.err
( time python bench.py --impl reference --steps 5 --warmup 3 > $out/${tag}_bench_reference.json ) 2> $out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_launches.log 2>&1
