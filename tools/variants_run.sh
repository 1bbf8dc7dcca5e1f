# compares the stock library with compile-time variants (tools/build_variant.sh <name> ...) on the headline bench
run() { env $1 python bench.py --steps 16 --warmup 3 --cpu-seconds 0.3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$2 ms=%.3f closest=%.3f shadow=%.3f shade=%.3f  Mrays/s=%.1f'%(d['ms_per_step'],k['trace_closest'],k['trace_shadow'],k['shade'],d['value']))"; }
run A=1 "stock        "
for v in ${VARIANTS}; do
  run LUMB200_LIBRARY=$PWD/luminary_b200/liblumb200_$v.so "variant $v"
done
run A=1 "stock again  "
