#!/bin/bash
# usage (under gpurun): bash tools/ncu_bench.sh <tag> [bench args...]
#   1. the plain program must exit 0 first; 2. launch list (gpu__time_duration.sum) -> gpurun_out/launches_<tag>.csv;
#   3. ncu --set full of two k_cost_volume launches -> gpurun_out/prof_<tag>.ncu-rep (read here with tools/ncu_summary.py)
tag=$1; shift
CMD="python bench.py --steps 2 --warmup 3 --batch 2 --no-e2e --no-cpu-baseline --no-bands $@"
$CMD > gpurun_out/plain_$tag.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_cost_volume -s 6 -c 2 -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
