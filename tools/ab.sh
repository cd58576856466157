#!/bin/bash
# A/B timing of library variants on ONE box: bash tools/ab.sh "<script and args>" variant1 variant2 ...   (ppodash_b200/libppd_<variant>.so)
CMD=$1; shift
for i in 1 2; do for v in "$@"; do
  echo -n "$v: "
  PPD_LIB=$PWD/ppodash_b200/libppd_$v.so timeout 200 python $CMD 2>&1 | python -c "
import sys,json
r=[json.loads(l) for l in sys.stdin if l.startswith('{')]
k='conv' if any('conv' in x for x in r) else 'gemm'
print(' '.join(f\"{x[k]}={x['ms']:.3f}\" for x in r if k in x), r[-1])"
done; done
