#!/bin/bash
# ncu launch list (gpu__time_duration + DRAM bytes) of one Barnes-Hut force evaluation:  scripts/prof_bh_launches.sh N IC TAG
n=${1:-1048576}; ic=${2:-plummer}; tag=${3:-bh}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_${tag}.csv python scripts/prof_bh.py $n 1 $ic > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_${tag}.csv > gpurun_out/r2_launches_${tag}_summary.csv
cat gpurun_out/r2_launches_${tag}_summary.csv
