#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one bench step with its device time, (2) --set full captures of the
# top kernels.  Run under gpurun from the repo root:  bash tools/profile_gpu.sh <tag>
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --cpu-sample 2000 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in mc_seed_kernel mc_locate_kernel mc_scatter_kernel mc_alnprep_kernel mc_alnfin_kernel mc_rescue_kernel mc_dp_kernel mc_cluster_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 1 -c 1 -o gpurun_out/full_${K}_${TAG} -f \
      python bench.py --steps 1 --warmup 1 --cpu-sample 2000 > /dev/null 2>&1
done
ls -la gpurun_out
