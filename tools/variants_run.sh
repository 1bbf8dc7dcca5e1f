# k_shade phase-split experiment: stock library unsplit / split, then variants with other occupancy targets (split)
run() { env $1 $2 python bench.py --steps 12 --warmup 3 --cpu-seconds 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$3 ms=%.3f closest=%.3f shadow=%.3f shade=%.3f  Mrays/s=%.1f'%(d['ms_per_step'],k['trace_closest'],k['trace_shadow'],k['shade'],d['value']))"; }
run A=1 LUMB200_SHADE_SPLIT=0 "stock unsplit      "
run A=1 LUMB200_SHADE_SPLIT=1 "stock split (5,5)  "
for v in s65 s66 s64 s44 s85; do
  run LUMB200_LIBRARY=$PWD/luminary_b200/liblumb200_$v.so LUMB200_SHADE_SPLIT=1 "variant $v split "
done
