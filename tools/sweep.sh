#!/bin/bash
# Sweeps the traversal tunables (env overrides) on the bench workload; prints ms/step and per-kernel ms.
for f in ${FETCH:-16 22 28}; do for t in ${TRI:-8 14 20}; do
  LUMB200_FETCH_THRESHOLD=$f LUMB200_TRI_THRESHOLD=$t python bench.py --steps 8 --warmup 3 ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('fetch=$f tri=$t ms=%.3f closest=%.3f shadow=%.3f shade=%.3f sort=%.3f'%(d['ms_per_step'],k['trace_closest'],k['trace_shadow'],k['shade'],k['sort']))"
done; done
