#!/bin/bash
# A/B helper (run under gpurun): parity subset for the fused binary rollout, bench at 20 and 128 steps per launch, and the
# executed / local-memory instruction counts of two steady-state 20-step launches.  Usage: tools/ab_rollout.sh <tag>
tag=${1:-ab}
cd "$(dirname "$0")/.."
timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "incremental or rollout_api or full_size or (batched_rollout_matches_oracle and (binary or zelda)) or step_host or graph_capturable" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu"
timeout 200 $B > gpurun_out/${tag}_T20.json 2>gpurun_out/${tag}.err
timeout 200 $B --steps 512 > gpurun_out/${tag}_T128.json 2>>gpurun_out/${tag}.err
for f in ${tag}_T20 ${tag}_T128; do python -c "
import json,sys; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', 'value %.4g closed_loop %.4g e2e %.4g e2e_rollout %.4g' % (d['value'], d['closed_loop']['value'], d['e2e']['value'], d.get('e2e_rollout',{}).get('value',0)))"; done
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum --clock-control none -k regex:k_rollout -s 8 -c 2 --csv --log-file gpurun_out/${tag}_inst.csv $B --repeats 5 --only-rollout > /dev/null 2>&1
tail -8 gpurun_out/${tag}_inst.csv | awk -F'","' '{print $(NF-2), $NF}'
