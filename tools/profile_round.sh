#!/bin/bash
# GPU box: launch list of the bench command, per-launch metrics of one step, one --set full capture of the top GEMM.
# usage: tools/profile_round.sh <tag>
tag=${1:-r01x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/step_metrics_${tag}.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    python tools/one_step.py --single-stream > gpurun_out/one_step_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16x3 -s 4 -c 3 -f -o gpurun_out/prof_gemm_${tag} \
    python tools/one_step.py --single-stream > gpurun_out/one_step_full_${tag}.log 2>&1
ls -la gpurun_out
