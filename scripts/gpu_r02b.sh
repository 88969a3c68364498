#!/bin/bash
# round 2, call b: the new full-size parity tests + ncu --set full of K2 / K3 / K4a on the default bench workload
out=gpurun_out; tag=r02b
mkdir -p $out
set -x
free -g | head -2 > $out/${tag}_host.txt; nproc >> $out/${tag}_host.txt
timeout 1700 python -m pytest tests/test_gpu_score.py tests/test_gpu_checkpoint.py tests/test_gpu_full_configs.py -q -m gpu --durations=12 > $out/${tag}_new_tests.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_new_tests.log
tail -25 $out/${tag}_new_tests.log
ncu --set full --clock-control none --import-source on -k regex:'k_enc_fwd|k_dec_fwd|k_sort' -s 30 -c 6 -f -o $out/${tag}_prof_k2k3k4a \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_k2k3.log 2>&1
ncu -i $out/${tag}_prof_k2k3k4a.ncu-rep --page raw --csv > $out/${tag}_prof_k2k3k4a_raw.csv 2>/dev/null
ls -la $out | tail
