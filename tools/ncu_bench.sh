#!/bin/bash
# usage (under gpurun): bash tools/ncu_bench.sh <tag> [bench args...]; writes gpurun_out/launches_<tag>.csv and prof_<tag>.ncu-rep
tag=$1; shift
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $@"
$CMD > gpurun_out/plain_$tag.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_cost_volume -s 8 -c 2 -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
