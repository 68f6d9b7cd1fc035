#!/bin/bash
# A/B: env warps per CTA (= per search arena) of k_rollout_async; variant libraries are built with -DASYNC_WPB=<w> into
# build/variants/ (git-ignored).  Usage (on the GPU box): bash tools/sweep_async_wpb.sh
for w in 4 1 2 8; do
  if [ $w = 4 ]; then L="X=1"; else L="PCGRL_B200_LIB=/root/repo/build/variants/libpcgrl_wpb$w.so"; fi
  for wl in sokoban-wide-5x5 sokoban-wide-5x5-sparse mdungeon-wide-default ddave-narrow-default; do
    env $L python bench.py --workload $wl --steps 256 --warmup 128 --only-rollout --no-cpu --no-sweep --no-flush-l2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('WPB=$w %-26s value %.3e' % ('$wl', d['value']))"
  done
done
