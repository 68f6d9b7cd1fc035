#!/bin/bash
# Round-2 final profiling pass (run under gpurun on ONE GPU) for the tree with the incremental binary statistics:
# launch list of the driver's bench command (durations + executed warp instructions per launch), the same for 128 steps
# per launch, and one ncu --set full capture of k_rollout<binary> at the driver's launch shape (T = 20, 4096 envs).
set -x
cd "$(dirname "$0")/.."
B="python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu --repeats 5"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02k_launches_bench_steps20.csv $B > gpurun_out/r02k_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_rollout -c 12 --csv --log-file gpurun_out/r02k_launches_rollout_T128.csv python bench.py --steps 512 --warmup 5 --no-sweep --no-cpu --repeats 2 --only-rollout > gpurun_out/r02k_ncu_bench128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 8 -c 1 -o gpurun_out/r02k_rollout_T20 -f $B --only-rollout > gpurun_out/r02k_ncu_full1.log 2>&1
ls -la gpurun_out/*.ncu-rep
