#!/bin/bash
# GPU box (end of round 2): per-launch metrics of one head step, the search's launch list at a 125k-row shard, and a --set full
# capture of the first-chunk selection kernel.
tag=${1:-r02s}
mkdir -p gpurun_out
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/step_metrics_${tag}.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    python tools/one_step.py --single-stream > gpurun_out/one_step_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/search_launches_${tag}.csv \
    python tools/time_search.py 125000 --once > gpurun_out/search_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:list_boot_select -s 1 -c 1 -f -o gpurun_out/prof_boot_${tag} \
    python tools/time_search.py 125000 --once >> gpurun_out/search_${tag}.log 2>&1
ls -la gpurun_out | tail -6
