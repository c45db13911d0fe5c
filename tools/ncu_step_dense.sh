#!/bin/bash
# The GEMM-relevant part of tools/ncu_step.sh alone (after a change to the dense kernels only): per-kernel sections of one engine
# step, the DRAM-traffic capture bench.py reads, and the launch list of a 2-step bench run.
set -x
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats \
    --section WarpStateStats --section ComputeWorkloadAnalysis --clock-control none --profile-from-start off -f -o /tmp/r02_step \
    python tools/profile_step.py > gpurun_out/r02_step_ncu.log 2>&1
python tools/ncu_metrics.py /tmp/r02_step.ncu-rep > gpurun_out/r02_step_metrics.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    -f -o /tmp/r02_traffic python tools/profile_step.py > gpurun_out/r02_traffic_ncu.log 2>&1
python tools/ncu_traffic.py /tmp/r02_traffic.ncu-rep gpurun_out/r02_traffic.json gpurun_out/r02_step_kernels.txt > gpurun_out/r02_traffic_summary.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --no-large --no-configs --no-sweep --no-dropin --no-cpu-baseline > gpurun_out/r02_launch_ncu.log 2>&1
head -12 gpurun_out/r02_traffic_summary.txt
