#!/bin/bash
# ncu evidence of one engine step, processed ON the GPU box (only text comes back; the .ncu-rep files stay in /tmp):
#   1. every kernel of the step with the sections the roofline needs (duration, DRAM bytes, occupancy, issue / warp stalls)
#      -> gpurun_out/r02_traffic.json, r02_step_kernels.txt, r02_step_metrics.txt
#   2. `--set full` for the kernels SURVEY 8(d) lists (bag_*, gine_*, segment_pool_*, ego_*) -> gpurun_out/r02_full_sparse_metrics.txt
#   3. the launch list of a 2-step bench run -> gpurun_out/r02_bench_launches_ncu.csv
set -x
python tools/profile_step.py > gpurun_out/r02_step_plain.log 2>&1 || exit 1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats \
    --section WarpStateStats --section ComputeWorkloadAnalysis --clock-control none --profile-from-start off -f -o /tmp/r02_step \
    python tools/profile_step.py > gpurun_out/r02_step_ncu.log 2>&1
python tools/ncu_traffic.py /tmp/r02_step.ncu-rep gpurun_out/r02_traffic.json gpurun_out/r02_step_kernels.txt > gpurun_out/r02_traffic_summary.txt 2>&1
python tools/ncu_metrics.py /tmp/r02_step.ncu-rep > gpurun_out/r02_step_metrics.txt 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k 'regex:bag_|gine_|segment_pool|ego_' -f -o /tmp/r02_full \
    python tools/profile_step.py > gpurun_out/r02_full_ncu.log 2>&1
python tools/ncu_metrics.py /tmp/r02_full.ncu-rep > gpurun_out/r02_full_sparse_metrics.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-large --no-configs --no-sweep --no-dropin --no-cpu-baseline > gpurun_out/r02_small.json 2> gpurun_out/r02_small.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --no-large --no-configs --no-sweep --no-dropin --no-cpu-baseline > gpurun_out/r02_launch_ncu.log 2>&1
# 4. the resistance-distance kernels alone on 8192 graphs (config 2: ZINC-shaped h=3; config 4: molhiv-shaped h=4 with loops):
#    headline metrics + hottest source lines of the cycle-space kernel -> gpurun_out/r02_rd_fast_cfg{2,4}_{metrics,lines}.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ego_rd -s 4 -c 2 -f -o /tmp/r02_rd2 python tools/profile_encode.py 2 > gpurun_out/r02_rd2_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ego_rd -s 8 -c 4 -f -o /tmp/r02_rd4 python tools/profile_encode.py 4 > gpurun_out/r02_rd4_ncu.log 2>&1
for c in 2 4; do
    python tools/ncu_metrics.py /tmp/r02_rd$c.ncu-rep > gpurun_out/r02_rd_fast_cfg${c}_metrics.txt 2>&1
    python tools/ncu_lines.py /tmp/r02_rd$c.ncu-rep ego_rd_fast 40 > gpurun_out/r02_rd_fast_cfg${c}_lines.txt 2>&1
done
ls -la gpurun_out /tmp/*.ncu-rep
