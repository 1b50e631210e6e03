#!/bin/bash
# Same-box A/B of libannp_b200 builds at the bench workload (kernel experiments; box-to-box variation is ~4 %, so only
# numbers from ONE call compare):  scripts/ab_bench.sh cur alt1 alt2 ...   ("cur" = meng_zhang_b200/lib, others = lib_alt/libannp_b200_<name>.so)
for v in "$@"; do
  if [ "$v" = cur ]; then L=""; else L=$PWD/meng_zhang_b200/lib_alt/libannp_b200_$v.so; fi
  ANNP_B200_LIB=$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-published-deck 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms_per_step %.3f kernel_ms %.3f frac %.3f e2e %.3e' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))"
done
