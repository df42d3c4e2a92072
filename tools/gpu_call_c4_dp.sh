#!/bin/bash
# configs[4] (ksw2, 250 bp, indel-rich): launch list of one batch and ncu --set full of the gapped-fill kernels (DP pipe evidence)
TAG=${1:-c4a}
mkdir -p gpurun_out
C="--config 4 --pairs 500000 --steps 1 --warmup 1 --resident-only --no-cpu --cache-dir /dev/shm/mc"
timeout 900 python bench.py $C > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err          # builds the index cache
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $C > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_${TAG}.csv | tee gpurun_out/launch_summary_${TAG}.txt | head -30
SKIP=$(python tools/launch_summary.py gpurun_out/launches_${TAG}.csv --skip-for "mc_(dp|dp_small|piece|alnprep|alnfin)_kernel")
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mc_(dp|dp_small|piece|alnprep|alnfin)_kernel" --launch-skip $SKIP -c 12 -o gpurun_out/full_${TAG} -f python bench.py $C > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/full_${TAG}.ncu-rep | tee gpurun_out/full_summary_${TAG}.txt
