#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one bench step with its device time, (2) --set full captures (with source
# correlation, needs -lineinfo) of the kernels named on the command line.  Run under gpurun from the repo root:
#   bash tools/profile_gpu.sh <tag> kernel[:skip] ...
TAG=${1:-r01}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --cpu-sample 2000 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for SPEC in "$@"; do
  K=${SPEC%%:*}; SKIP=${SPEC##*:}; [ "$SKIP" = "$SPEC" ] && SKIP=1
  ncu --set full --clock-control none --import-source on -k regex:${K} -s ${SKIP} -c 1 -o gpurun_out/full_${K}_${TAG} -f \
      python bench.py --steps 1 --warmup 1 --cpu-sample 2000 > /dev/null 2>&1
done
ls -la gpurun_out | tail -12
