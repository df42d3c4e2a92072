#!/bin/bash
# GRCh38-sized (configs[3] / configs[4]) evidence in one gpurun call: e2e trace, launch list, --set full of the large kernels.
# usage: bash tools/gpu_call_c3.sh <tag> <config> [steps...]   steps: trace launches full
TAG=${1:-c3a}; CFGN=${2:-3}; shift; shift
STEPS=${@:-trace launches full}
mkdir -p gpurun_out
CACHE="--cache-dir /dev/shm/mc"   # the first step builds the index, the runs under ncu load it
for S in $STEPS; do
case $S in
trace)
  MC_VC_TRACE=1 MC_DEBUG=1 timeout 900 python bench.py --config $CFGN --pairs 4000000 --steps 1 --warmup 1 --no-cpu --no-sam --trace-e2e $CACHE > gpurun_out/trace_${TAG}.json 2> gpurun_out/trace_${TAG}.err
  grep -E "e2e vcf|variant_scan|vc\]" gpurun_out/trace_${TAG}.err | tail -80 ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --config $CFGN --pairs 2000000 --steps 1 --warmup 1 --resident-only --no-cpu $CACHE > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_${TAG}.csv | tee gpurun_out/launch_summary_${TAG}.txt | head -40 ;;
full)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mc_(seed|locate|cluster|pair|rescue|alnprep|alnfin|scatter)_kernel" \
      --launch-skip ${FULL_SKIP:-8} -c ${FULL_COUNT:-14} -o gpurun_out/full_${TAG} -f \
      python bench.py --config $CFGN --pairs 2000000 --steps 1 --warmup 1 --resident-only --no-cpu $CACHE > gpurun_out/full_${TAG}.log 2>&1
  ls -la gpurun_out/full_${TAG}.ncu-rep ;;
esac
done
