#!/bin/bash
# A/B of library variants (profiles/tools/build_variant.sh) on the headline MCTS leg: one line per variant
#   profiles/tools/ab_variants.sh OUT.log name1 name2 ...   ("base" = the shipped library); extra bench flags in $BENCH_FLAGS
out=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset SPL_B200_LIB; else export SPL_B200_LIB=$PWD/alphazero-general-ori_b200/build/variants/libsplendor_b200_$v.so; fi
  timeout 300 python bench.py --workload mcts --no-cpu --no-extra --wide-trees 0 --steps 4 $BENCH_FLAGS 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=l['roofline']['wave_breakdown_ms']
print('$v', 'sims/s %.3e' % l['value'], 'e2e %.3e' % l['e2e']['value'], 'wave_us %.1f' % (1e3*l['roofline']['avg_launch_ms']), 'sel %.1f nn %.1f exp %.1f' % tuple(1e3*x for x in list(b.values())[:3]), flush=True)
" >> $out 2>&1
done
