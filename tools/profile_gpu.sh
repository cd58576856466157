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
# (1) launch list of one stretch of the timed region: every kernel with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list rc=$? (skipped $SKIP)"
# (2) DRAM traffic of the dominant kernel family over 64 consecutive launches (= 2 minibatches of 32 tca launches)
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tca_gemm_kernel -s 200 -c 64 --csv --log-file $OUT/tca_traffic_$TAG.csv $CMD > $OUT/ncu_t_$TAG.log 2>&1
echo "tca traffic rc=$?"
# (3) full captures of the top kernels (the 20th tca launch is a convolution)
for K in tca_gemm_kernel gru_fwd_cluster512 gru_bwd_cluster512; do
  $CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 1 -o $OUT/prof_${K}_$TAG -f $CMD > $OUT/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la $OUT | grep $TAG
