#!/bin/bash
# One gpurun call of round 2: GPU tests, the bench line, the launch list and --set full captures of the large kernels at
# the bench workload (configs[2]).  usage: bash tools/gpu_call.sh <tag> [steps...]   steps: tests bench trace launches full
TAG=${1:-r2a}; shift
STEPS=${@:-tests bench trace launches full}
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
for S in $STEPS; do
case $S in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/tests_${TAG}.log ;;
bench)
  timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 6000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err ;;
trace)
  MC_VC_TRACE=1 MC_DEBUG=1 timeout 600 python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu > gpurun_out/trace_${TAG}.json 2> gpurun_out/trace_${TAG}.err; tail -60 gpurun_out/trace_${TAG}.err ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --pairs 2000000 --steps 1 --warmup 1 --resident-only --no-cpu > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_${TAG}.csv | tee gpurun_out/launch_summary_${TAG}.txt | head -40 ;;
full)
  # the second pass over the batch (launch-skip past the first, cold pass): every distinct large kernel once
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mc_(seed|locate|cluster|pair|alnprep|piece|dp_small|dp|alnfin|profkey|scatter|profpiece|rescue)_kernel" \
      --launch-skip ${FULL_SKIP:-27} -c ${FULL_COUNT:-30} -o gpurun_out/full_${TAG} -f \
      python bench.py --pairs 2000000 --steps 1 --warmup 1 --resident-only --no-cpu > gpurun_out/full_${TAG}.log 2>&1
  ls -la gpurun_out/full_${TAG}.ncu-rep ;;
esac
done
