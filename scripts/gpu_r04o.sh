#!/bin/bash
# round 2, session 3, call O: batch-side streams (fill, K1, K4a) at the lowest priority (OCF_GATHER_PRIO=0) against the highest
out=gpurun_out; tag=r04o; mkdir -p $out
B="python bench.py --others none --no-cpu-baseline --no-scoring"
OCF_GATHER_PRIO=0 timeout 60 $B > $out/${tag}_ml10m_lowprio.json 2> $out/${tag}_ml10m_lowprio.err
timeout 60 $B > $out/${tag}_ml10m_hiprio.json 2> $out/${tag}_ml10m_hiprio.err
python scripts/show_line.py $out/${tag}_ml10m_lowprio.json $out/${tag}_ml10m_hiprio.json
