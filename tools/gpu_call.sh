#!/bin/bash
# One gpurun call of round 2: GPU tests, the bench line, the launch list and --set full captures of the large kernels at
# the bench workload (configs[2]; EXTRA="--config 3 --cache-dir /dev/shm/mc" for the GRCh38-sized one).
# usage: bash tools/gpu_call.sh <tag> [steps...]   steps: tests bench trace launches full
TAG=${1:-r2a}; shift
STEPS=${@:-tests bench trace launches full}
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
for S in $STEPS; do
case $S in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/tests_${TAG}.log ;;
bench)
  timeout 900 python bench.py $EXTRA > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 6000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err ;;
trace)
  MC_VC_TRACE=1 MC_DEBUG=1 timeout 600 python bench.py --pairs 2000000 --steps 1 --warmup 1 --no-cpu > gpurun_out/trace_${TAG}.json 2> gpurun_out/trace_${TAG}.err; tail -60 gpurun_out/trace_${TAG}.err ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --pairs 2000000 --steps 1 --warmup 1 --resident-only --no-cpu $EXTRA > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_${TAG}.csv | tee gpurun_out/launch_summary_${TAG}.txt | head -40 ;;
full)
  # the timed pass over the batch (the launches before it are the warm-up pass): every distinct large kernel once.  The number
  # of matching launches to skip comes from the launch list of the same command (step `launches`).
  RE="mc_(seed|locate|cluster|pair|rescue|alnprep|piece|dp_small|dp|alnfin|profkey|scatter|profpiece)_kernel"
  SKIP=$(python tools/launch_summary.py gpurun_out/launches_${TAG}.csv --skip-for "$RE" 2>/dev/null || echo ${FULL_SKIP:-27})
  echo "ncu --set full: skipping $SKIP matching launches"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" \
      --launch-skip $SKIP -c ${FULL_COUNT:-30} -o gpurun_out/full_${TAG} -f \
      python bench.py --pairs 2000000 --steps 1 --warmup 1 --resident-only --no-cpu $EXTRA > gpurun_out/full_${TAG}.log 2>&1
  ls -la gpurun_out/full_${TAG}.ncu-rep
  python tools/ncu_summary.py gpurun_out/full_${TAG}.ncu-rep | tee gpurun_out/full_summary_${TAG}.txt ;;
esac
done
