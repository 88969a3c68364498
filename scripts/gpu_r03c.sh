#!/bin/bash
# round 2 evidence on the final code: the driver's own bench command, ncu --set full of the step's kernels (ML-10M shape)
# and of the tcgen05 hidden-layer contraction (ML-20M shape), launch lists
out=gpurun_out; tag=${1:-r03c}
mkdir -p $out
( time python bench.py > $out/${tag}_bench_default.json ) 2> $out/${tag}_bench_default.err
tail -c 300 $out/${tag}_bench_default.json; tail -4 $out/${tag}_bench_default.err
python bench.py --impl reference --steps 5 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
ncu --set full --clock-control none --import-source on -k regex:'k_enc_fwd|k_dec_fwd|k_sort|k_row_update|k_dz_bias|k_gather' -s 60 -c 9 -f -o $out/${tag}_prof_step_ml10m \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_step.log 2>&1
ncu -i $out/${tag}_prof_step_ml10m.ncu-rep --page raw --csv > $out/${tag}_prof_step_ml10m_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc' -s 30 -c 3 -f -o $out/${tag}_prof_gemm_tc_ml20m \
    python bench.py --workload ml20m --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_gemm.log 2>&1
ncu -i $out/${tag}_prof_gemm_tc_ml20m.ncu-rep --page raw --csv > $out/${tag}_prof_gemm_tc_ml20m_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_ml10m.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --others none --no-scoring > $out/${tag}_ncu_launches.log 2>&1
ls -la $out | grep ${tag}
