#!/bin/bash
# One GPU-box call, many measurements: every step has its own timeout and log under gpurun_out/batch/, a failing
# step does not stop the rest.  A box call costs 1-3.5 GPU-minutes of overhead whatever it runs, so batch.
#   usage (on the GPU box, from the repo root):  bash tools/gpu_batch.sh [tests] [newtests] [bench] [kernels] [probes] [softmax] [cross] [ncu]
set -u
out=gpurun_out/batch; mkdir -p "$out"
step() {  # step <name> <timeout-seconds> <command...>
  local name=$1 to=$2; shift 2
  local t0=$SECONDS
  timeout "$to" "$@" > "$out/$name.log" 2>&1
  echo "[$name] rc=$? $((SECONDS - t0))s  $(tail -1 "$out/$name.log" | cut -c1-200)"
}
[ $# -eq 0 ] && set -- tests bench kernels
for what in "$@"; do
  case $what in
    tests)    step pytest_gpu 900 python -m pytest tests -q -m gpu -p no:cacheprovider ;;
    newtests) step pytest_new 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_tc.py -q -m gpu -p no:cacheprovider ;;
    newtests2) step pytest_new2 900 python -m pytest tests/test_gpu_round2b.py -q -m gpu -p no:cacheprovider --timeout 180 ;;
    ncukern)  step ncu_kern_launches 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_kernels.csv" python tools/kernel_step_probe.py 2 ;;
    ncukfull) step ncu_kern_full 600 ncu --set full --clock-control none --import-source on -k regex:"gather_warpchunk|ag_|cross_tc_kernel" -s 13 -c 13 -f -o "$out/prof_kernels" python tools/kernel_step_probe.py 2 ;;
    ncugather) step ncu_gather 300 ncu --set full --clock-control none --import-source on -k regex:"gather_warpchunk" -c 4 -f -o "$out/prof_gather" python tools/kernel_step_probe.py 2 ;;
    ncusel)   step ncu_sel 300 ncu --set full --clock-control none --import-source on -k regex:"tc_select|tc_rescore|tc_threshold" -s 9 -c 3 -f -o "$out/prof_finalize" python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline ;;
    crossalign) step cross_align 120 python tools/cross_align_probe.py ;;
    multi)    step pytest_multi 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider ;;
    bench)    step bench_n1 300 python bench.py --steps 30 --warmup 5 ;;
    benchq)   step bench_n1_quick 200 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline ;;
    ref)      step bench_ref 300 python bench.py --impl reference --steps 5 --warmup 1 ;;
    kernels)  step bench_kernels 200 python tools/bench_kernels.py ;;
    probes)   step tc_rate_probe 60 python tools/tc_rate_probe.py; step hbm_probe 100 python tools/hbm_probe.py ;;
    softmax)  step softmax_launches 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_softmax.csv" python tools/softmax_probe.py 16384 64 2 ;;
    cross)    step cross_probe 200 python tools/cross_probe.py; step cross_launches 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_cross.csv" python tools/cross_step_probe.py 2 ;;
    ncu)      step ncu_launches 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline ;;
    ncufin)   step ncu_fin 300 ncu --set full --clock-control none --import-source on -k regex:tc_finalize -s 3 -c 1 -f -o "$out/prof_finalize" python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline ;;
    ncuscan)  step ncu_scan 300 ncu --set full --clock-control none --import-source on -k regex:tc_scan -s 6 -c 2 -f -o "$out/prof_scan" python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline ;;
    adagrad)  step adagrad_launches 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_adagrad.csv" python tools/adagrad_probe.py 2; step adagrad_probe 60 python tools/adagrad_probe.py 20 ;;
    benchn)   step bench_multi 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-2} --steps 30 --warmup 5 ;;
    benchnccl) TFRS_SHARD_EXCHANGE=nccl step bench_multi_nccl 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus ${NGPU:-2} --steps 30 --warmup 5 ;;
    benchn4)  step bench_multi_cfg4 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${NGPU:-2} --steps 10 --warmup 3 --workload cfg4 ;;
    *) echo "unknown step $what" ;;
  esac
done
