#!/bin/bash
# Run under gpurun:  bash tools/profile_gpu.sh <tag>
# Each ncu run is preceded by the same command without ncu (B200_PROFILING.md); outputs land in gpurun_out/.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-micro --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
N=$(python -c "import json;print(json.load(open('$OUT/plain_$TAG.log'))['gpu_launches'])" 2>/dev/null || echo 3500)
SKIP=$((N + N / 4 + 200))
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list rc=$? (skipped $SKIP)"
for K in tc_gemm_kernel gru_fwd_cluster512 gru_bwd_cluster512; do
  $CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 1 -o $OUT/prof_${K}_$TAG -f $CMD > $OUT/ncu_$K_$TAG.log 2>&1
  echo "$K rc=$?"
done
python tools/microbench.py returns > $OUT/mb_returns_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:returns_scan -s 30 -c 1 -o $OUT/prof_returns_scan_$TAG -f python tools/microbench.py returns > $OUT/ncu_ret_$TAG.log 2>&1
echo "returns rc=$?"
python tools/microbench.py gather > $OUT/mb_gather_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gather_rows -s 90 -c 1 -o $OUT/prof_gather_rows_$TAG -f python tools/microbench.py gather > $OUT/ncu_gat_$TAG.log 2>&1
echo "gather rc=$?"
ls -la $OUT | grep $TAG
