#!/bin/bash
for np in 1 2 4 6 8; do
  CRB_RING_PRODUCERS=$np CRB_BPR_RING=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --eval-users 0 --no-eval-full --optimizer SGD 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('producers $np SGD ms/step %.3f K3 %.3f'%(d['ms_per_step'], r['kernel_ms']))"
done
