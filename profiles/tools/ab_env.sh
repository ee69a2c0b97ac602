#!/bin/bash
# A/B of library variants on the env leg: profiles/tools/ab_env.sh OUT.log name1 name2 ...
out=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset SPL_B200_LIB; else export SPL_B200_LIB=$PWD/alphazero-general-ori_b200/build/variants/libsplendor_b200_$v.so; fi
  timeout 300 python bench.py --workload env --no-cpu --no-extra --no-sweep $BENCH_FLAGS 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'env %.4e steps/s frac %.3f e2e %.4e' % (l['value'], l['roofline']['frac'], l['e2e']['value']), flush=True)" >> $out 2>&1
done
