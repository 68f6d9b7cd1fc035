for n in 4096 8192 16384 32768 65536; do
 for p in 1 0; do
  PCGRL_PACKED=$p python bench.py --steps 256 --warmup 128 --no-sweep --no-cpu --only-rollout --envs $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n=$n packed=$p value %.3e frac %.3f launch_ms %.3f' % (d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))"
 done
done
