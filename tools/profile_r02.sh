#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE GPU): launch list of the bench command, ncu --set full captures of the
# headline rollout kernel at the driver's launch shape (T=20, 4096 envs), of the single-step launch (T=1) and of the smb
# rollout kernel.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
set -x
cd "$(dirname "$0")/.."
B="python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu --repeats 5"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench_steps20.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 5 -c 2 -o gpurun_out/r02_rollout_T20 -f $B --only-rollout > gpurun_out/r02_ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 60 -c 2 -o gpurun_out/r02_rollout_T1 -f $B > gpurun_out/r02_ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_smb_rollout -s 1 -c 1 -o gpurun_out/r02_smb_rollout -f python tools/bench_smb.py --envs 512 --steps 4 > gpurun_out/r02_ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
