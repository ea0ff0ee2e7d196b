#!/bin/bash
# One GPU-box call, many measurements: every step has its own timeout and log under gpurun_out/batch/, a failing
# step does not stop the rest.  A box call costs 1-3.5 GPU-minutes of overhead whatever it runs, so batch.
#   usage (on the GPU box, from the repo root):  bash tools/gpu_batch.sh [tests] [bench] [kernels] [probes] [softmax] [cross]
set -u
out=gpurun_out/batch; mkdir -p "$out"
step() {  # step <name> <timeout-seconds> <command...>
  local name=$1 to=$2; shift 2
  local t0=$SECONDS
  timeout "$to" "$@" > "$out/$name.log" 2>&1
  echo "[$name] rc=$? $((SECONDS - t0))s  $(tail -1 "$out/$name.log" | cut -c1-160)"
}
[ $# -eq 0 ] && set -- tests bench kernels
for what in "$@"; do
  case $what in
    tests)   step pytest_gpu 400 python -m pytest tests -q -m gpu ;;
    bench)   step bench_n1 200 python bench.py --steps 30 --warmup 5 ;;
    kernels) step bench_kernels 200 python tools/bench_kernels.py ;;
    probes)  step tc_rate_probe 60 python tools/tc_rate_probe.py; step hbm_probe 100 python tools/hbm_probe.py; step gather_probe 100 python tools/gather_probe.py ;;
    softmax) step softmax_launches 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_softmax.csv" python tools/softmax_probe.py 16384 64 2 ;;
    cross)   step cross_probe 200 python tools/cross_probe.py; step cross_launches 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cross.csv" python tools/cross_step_probe.py 2 ;;
    *) echo "unknown step $what" ;;
  esac
done
