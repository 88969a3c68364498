#!/bin/bash
# round 2 final evidence on the final code (one box): the full GPU suite, the driver's own bench command, ncu --set full of
# every kernel of the headline step + the Netflix-shape row update, launch lists of three workloads
out=gpurun_out; tag=${1:-r04z}
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
tail -3 $out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
( time python bench.py > $out/${tag}_bench_default.json ) 2> $out/${tag}_bench_default.err
tail -c 300 $out/${tag}_bench_default.json; tail -4 $out/${tag}_bench_default.err
python bench.py --steps 20 --warmup 3 --others none --no-cpu-baseline --no-scoring > $out/${tag}_bench_steps20.json 2> $out/${tag}_bench_steps20.err
ncu --set full --clock-control none --import-source on -k regex:'k_enc_fwd|k_dec_fwd|k_sort|k_row_update|k_dz_bias|k_gather' -s 60 -c 16 -f -o $out/${tag}_prof_step_ml10m \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_step.log 2>&1
ncu -i $out/${tag}_prof_step_ml10m.ncu-rep --page raw --csv > $out/${tag}_prof_step_ml10m_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_ml10m.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_launches.log 2>&1
for w in jester ml20m; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_$w.csv \
      python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_launches_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'k_row_update' -s 8 -c 1 -f -o $out/${tag}_prof_rowupd_netflix \
    python bench.py --workload netflix --steps 3 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_netflix.log 2>&1
ncu -i $out/${tag}_prof_rowupd_netflix.ncu-rep --page raw --csv > $out/${tag}_prof_rowupd_netflix_raw.csv 2>/dev/null
rm -f $out/${tag}_prof_step_ml10m.ncu-rep $out/${tag}_prof_rowupd_netflix.ncu-rep
ls -la $out | grep ${tag}
