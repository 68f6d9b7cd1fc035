#!/bin/bash
# A/B on the GPU box: default library (open-list top in shared memory, tail in HBM; 2 env warps x 4 arenas per SM) against
# build/variants/libpcgrl_allsmem.so (-DASYNC_HEAP_FAST=0 -DASYNC_WPB=4 -DASYNC_MIN_CTAS=2: whole open list in shared
# memory, 4 env warps x 2 arenas per SM).  Build the variant first (here, nvcc cross-compiles without a GPU):
#   mkdir -p build/variants && (cd gym_pcgrl_b200/csrc && nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false \
#     -shared -Xcompiler -fPIC -Xcompiler -pthread -DASYNC_HEAP_FAST=0 -DASYNC_WPB=4 -DASYNC_MIN_CTAS=2 \
#     -o ../../build/variants/libpcgrl_allsmem.so pcgrl_b200.cu pcgrl_linear.cu)
# Result (profiles/r02_summary.md): neutral within noise on every workload.
for L in "X=1" "PCGRL_B200_LIB=/root/repo/build/variants/libpcgrl_allsmem.so"; do
  echo "== $L"
  env $L python tools/bench_step_batch.py --workloads sokoban-wide-5x5-sparse,mdungeon-narrow-default,ddave-narrow-default --envs 8192,131072 --steps 32 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-26s %6d envs  device_step %.3e  e2e %.3e  %.2f ms/step' % (d['workload'], d['envs'], d['device_step'], d['e2e'], d['ms_per_step']))"
  for wl in sokoban-wide-5x5 sokoban-wide-5x5-sparse mdungeon-wide-default ddave-narrow-default; do
    env $L python bench.py --workload $wl --steps 256 --warmup 128 --only-rollout --no-cpu --no-sweep --no-flush-l2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rollout %-26s value %.3e' % ('$wl', d['value']))"
  done
done
