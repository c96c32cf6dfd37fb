#!/bin/bash
# GPU box (round 2): per-launch metrics of one head step, launch list of the bench command, --set full captures of the top kernels.
tag=${1:-r02m}
mkdir -p gpurun_out
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/step_metrics_${tag}.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    python tools/one_step.py --single-stream > gpurun_out/one_step_${tag}.log 2>&1
# the batched f2 forward GEMM (pair kernel, split-bf16) and one fp16 x1 gradient GEMM (pair kernel, single plane)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_pair_bf16x3 -s 3 -c 1 -f -o gpurun_out/prof_pairgemm_x3_${tag} \
    python tools/one_step.py --single-stream > gpurun_out/one_step_full1_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_pair_bf16x3 -s 60 -c 2 -f -o gpurun_out/prof_pairgemm_x1_${tag} \
    python tools/one_step.py --single-stream > gpurun_out/one_step_full2_${tag}.log 2>&1
# the search: coarse GEMM (steady chunk), first-chunk selection, list merge, re-score
ncu --set full --clock-control none --import-source on -k regex:"list_boot_select|list_update_warp|rescore_keys" -s 3 -c 3 -f -o gpurun_out/prof_search_${tag} \
    python tools/time_search.py 125000 --once > gpurun_out/search_full_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:coarse_gemm2 -s 20 -c 1 -f -o gpurun_out/prof_coarse_${tag} \
    python tools/time_search.py 125000 --once >> gpurun_out/search_full_${tag}.log 2>&1
ls -la gpurun_out | tail -12
